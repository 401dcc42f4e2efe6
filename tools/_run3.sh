set -x
python -m pytest tests -m gpu -q -x -k "fp32_mode_goldens or threshold or full_reference or large_batch" 2>&1 | tail -15 > gpurun_out/r02_t3.log
python tools/e2e_sweep.py > gpurun_out/r02_e2e_sweep.log 2>&1
SPART_HOST_THREADS=16 python tools/e2e_sweep.py 2>&1 | grep pageable > gpurun_out/r02_e2e_sweep_t16.log
cat gpurun_out/r02_t3.log gpurun_out/r02_e2e_sweep.log gpurun_out/r02_e2e_sweep_t16.log | grep -v "^{" 
