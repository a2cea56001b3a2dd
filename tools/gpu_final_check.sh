#!/bin/bash
# Run under gpurun (1 GPU): smoke(), the whole GPU test suite, the default bench line.
set -x
mkdir -p gpurun_out
TAG=${1:-r1l}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -2 gpurun_out/smoke_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > gpurun_out/gpu_tests_${TAG}.log
tail -12 gpurun_out/gpu_tests_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
