#!/bin/bash
# Run under gpurun (1 GPU).  Writes ncu artefacts to gpurun_out/; summaries are copied to profiles/ by
# tools/ncu_summary.py.   usage: tools/profile.sh TAG SCALE "kernel list"
set -x
mkdir -p gpurun_out
TAG=${1:-r1}
SCALE=${2:-0.2}
KERNELS=${3:-"fb_kernel env_kernel msv_kernel mdtrace_kernel mdclust_kernel"}
# 1) every launch of one whole step with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 3000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --scale $SCALE --steps 1 --warmup 3 --no-cpu-baseline --no-cli \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# 2) full-set capture of the DP kernels (few launches each; ~40 replays per launch)
for K in $KERNELS; do
  SKIP=12; [ "$K" = "msv_kernel" ] && SKIP=2; case "$K" in md*) SKIP=3;; esac
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 2 -f \
      -o gpurun_out/prof_${K}_${TAG} python bench.py --scale $SCALE --steps 1 --warmup 3 --no-cpu-baseline --no-cli \
      > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
ls -la gpurun_out
