#!/bin/bash
# Run under gpurun (one GPU).  Writes the evidence the roofline numbers come from to gpurun_out/:
#   <tag>_bench.json              the bench line (no profiler attached)
#   <tag>_launches.csv            every launch of our kernels in a short run of the same command, with device time
#   <tag>_counts_<cfg>_<prec>.csv executed FP64 / FP32 / SFU instructions and DRAM bytes per launch, per config
#   <tag>_prof_*.ncu-rep          ncu --set full captures (source-level) of every kernel family
# usage: tools/collect_profiles.sh <tag> [quick]
set -u
TAG=${1:-r02}
QUICK=${2:-}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { echo "plain bench failed"; tail -5 $OUT/${TAG}_bench.err; exit 1; }
KREG='regex:lidf_kernel|geometry_kernel|band_kernel|spectrum_kernel|lut_nearest'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 120 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/${TAG}_ncu_launches.log 2>&1
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
M=$M,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum
M=$M,smsp__inst_executed_pipe_xu.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for spec in "2 fp64" "2 fp32" "3 fp64" "3 fp32" "4 fp64" "4 fp32" "5 fp64" "5 fp32"; do
  set -- $spec
  python tools/profile_step.py --config $1 --precision $2 > /dev/null 2>&1 && \
  ncu --metrics $M --clock-control none --profile-from-start off -k "$KREG" --csv \
      --log-file $OUT/${TAG}_counts_cfg$1_$2.csv python tools/profile_step.py --config $1 --precision $2 > $OUT/${TAG}_ncu_counts.log 2>&1
done
[ -n "$QUICK" ] && exit 0
full() {  # name, then the profile_step arguments
  local name=$1; shift
  python tools/profile_step.py "$@" > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -k "$KREG" \
      -f -o $OUT/${TAG}_prof_$name python tools/profile_step.py "$@" > $OUT/${TAG}_ncu_full_$name.log 2>&1
  tail -1 $OUT/${TAG}_ncu_full_$name.log
}
full cfg2 --config 2 --n 1000000
full cfg3 --config 3 --n 1000000
full cfg2_fp32 --config 2 --precision fp32 --n 1000000
full srf --config 2 --mode srf --n 131072
full spectrum --config 2 --mode spectrum
full lut --config 2 --mode lut
ls -la $OUT/${TAG}_prof_*.ncu-rep
