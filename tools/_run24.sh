python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench.err
python tools/e2e_sweep.py > gpurun_out/r02_e2e_sweep.log 2>&1; grep pageable gpurun_out/r02_e2e_sweep.log | grep -v "^{"
python -m pytest tests -m gpu -q -x -k "host or pageable or f32_io or compact or broadcast" 2>&1 | tail -3
