set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_t12.log
cat gpurun_out/r02_t12.log
python tools/kbench.py 1000000 LANDSAT8-OLI 3 > gpurun_out/r02_kbench3.log 2>&1
python tools/kbench.py 1000000 LANDSAT8-OLI 3 fp32 >> gpurun_out/r02_kbench3.log 2>&1
python tools/kbench.py >> gpurun_out/r02_kbench3.log 2>&1
grep '^{' gpurun_out/r02_kbench3.log
