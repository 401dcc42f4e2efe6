set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r02_t11.log
cat gpurun_out/r02_t11.log
python __graft_entry__.py smoke 2>&1 | tail -4
( time python bench.py ) > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -5 gpurun_out/r02c_bench.err
( time python bench.py --impl reference ) > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err
tail -5 gpurun_out/r02c_bench_reference.err
