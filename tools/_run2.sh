set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_t2.log
bash tools/collect_profiles.sh r02a quick 2>&1 | tail -5
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err
cat gpurun_out/r02_t2.log
tail -3 gpurun_out/r02a_bench.err
