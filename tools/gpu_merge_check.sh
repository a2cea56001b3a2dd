set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_merge.py -x -q -s 2>&1 | tail -40 > gpurun_out/merge_tests.log
cat gpurun_out/merge_tests.log
timeout 300 python tools/merge_bench.py > gpurun_out/merge_bench.json 2> gpurun_out/merge_bench.err
cat gpurun_out/merge_bench.json; tail -5 gpurun_out/merge_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:merge_kernel -s 2 -c 1 -f -o gpurun_out/prof_merge_kernel_r1g python tools/merge_bench.py --pairs 200000 --steps 1 --warmup 2 --no-cpu > gpurun_out/ncu_merge.log 2>&1
tail -3 gpurun_out/ncu_merge.log
