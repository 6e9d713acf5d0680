#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c9_gpu_tests.log 2>&1
timeout 600 python tools/perf_probe.py 512 67108864 40 slab 0:0 > gpurun_out/c9_ab_64M.log 2>&1
MPM_B200_REORDER_PERIOD=0 timeout 600 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/c9_ab_64M.log 2>&1
MPM_PROBE_SHUFFLE=1 timeout 600 python tools/perf_probe.py 256 8388608 40 slab 0:0 >> gpurun_out/c9_ab_64M.log 2>&1
tail -3 gpurun_out/c9_gpu_tests.log; cat gpurun_out/c9_ab_64M.log
