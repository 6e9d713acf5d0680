#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c3_gpu_tests.log 2>&1
echo "default gpu tests: exit $?" >> gpurun_out/c3_gpu_tests.log
timeout 600 env MPM_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/c3_exp_tests.log 2>&1
echo "exp gpu tests: exit $?" >> gpurun_out/c3_exp_tests.log
timeout 600 python tools/perf_probe.py 512 67108864 10 slab 0:0,2:0,2:4,4:4 > gpurun_out/c3_ab_64M.log 2>&1
tail -n 4 gpurun_out/c3_gpu_tests.log gpurun_out/c3_exp_tests.log
cat gpurun_out/c3_ab_64M.log
