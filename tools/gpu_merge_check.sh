#!/bin/bash
# Run under gpurun (1 GPU): GPU test suite, merge-stage bench + ncu capture of merge_kernel, default bench line.
set -x
mkdir -p gpurun_out
TAG=${1:-r1h}
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30 > gpurun_out/gpu_tests_${TAG}.log
cat gpurun_out/gpu_tests_${TAG}.log
timeout 300 python tools/merge_bench.py > gpurun_out/merge_bench_${TAG}.json 2> gpurun_out/merge_bench.err
cat gpurun_out/merge_bench_${TAG}.json; tail -5 gpurun_out/merge_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:merge_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_merge_kernel_${TAG} python tools/merge_bench.py --pairs 200000 --steps 1 --warmup 2 --no-cpu \
    > gpurun_out/ncu_merge.log 2>&1
tail -2 gpurun_out/ncu_merge.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
