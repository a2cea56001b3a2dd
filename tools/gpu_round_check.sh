#!/bin/bash
# Run under gpurun (1 GPU): whole GPU test suite, paired command-line end-to-end timings, default bench line.
set -x
mkdir -p gpurun_out
TAG=${1:-r1j}
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > gpurun_out/gpu_tests_${TAG}.log
cat gpurun_out/gpu_tests_${TAG}.log
timeout 600 python tools/cli_e2e.py --paired unmerged > gpurun_out/cli_paired_unmerged_${TAG}.log 2>&1
head -1 gpurun_out/cli_paired_unmerged_${TAG}.log
timeout 600 python tools/cli_e2e.py --paired merged > gpurun_out/cli_paired_merged_${TAG}.log 2>&1
head -1 gpurun_out/cli_paired_merged_${TAG}.log


