set -x
python tools/kbench.py > gpurun_out/r02_kbench4.log 2>&1
python tools/kbench.py 1000000 LANDSAT8-OLI 3 >> gpurun_out/r02_kbench4.log 2>&1
grep '^{' gpurun_out/r02_kbench4.log || tail -20 gpurun_out/r02_kbench4.log
python -m pytest tests -m gpu -q -x -k "leafangles or lidf or large_batch or golden or edge" 2>&1 | tail -6
