#!/bin/bash
# ncu artefacts of the final round-2 code (run under gpurun, one GPU): launch list of one step + full captures.
TAG=${1:-r3m}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cli > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for K in fb_kernel fbdec_kernel mdclust_kernel; do
  SKIP=12; case "$K" in md*) SKIP=3;; esac
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:^$K -s $SKIP -c 1 -f \
      -o gpurun_out/prof_${K}_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cli \
      > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:deflate_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_deflate_kernel_${TAG} python tools/gzip_rate.py --mb 128 > gpurun_out/gzip_rate_under_ncu_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
