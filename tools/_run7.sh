set -x
bash tools/collect_profiles.sh r02 2>&1 | tail -12
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -2 gpurun_out/r02_bench.err
