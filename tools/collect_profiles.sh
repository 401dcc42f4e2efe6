#!/bin/bash
# Run under gpurun (one GPU).  Writes the evidence the roofline numbers come from to gpurun_out/:
#   bench_plain.json        the bench line (no profiler attached)
#   launches.csv            every kernel launch of the same command with its device time
#   prof.ncu-rep            ncu --set full capture of one launch of each of our kernels
# usage: tools/collect_profiles.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lidf_kernel|geometry_kernel|band_kernel|fma_chain" -c 60 \
    --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"lidf_kernel|geometry_kernel|band_kernel" -s 9 -c 3 \
    -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
