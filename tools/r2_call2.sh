#!/bin/bash
# Round 2, GPU call 2: tolerance-form F-update A/B + default suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/c2_gpu_tests.log 2>&1
echo "default gpu tests: exit $?" >> gpurun_out/c2_gpu_tests.log
timeout 600 python tools/perf_probe.py 512 67108864 10 slab 0:0,2:4,4:4,3:4 > gpurun_out/c2_ab_64M.log 2>&1
MPM_PROBE_FUPDATE_EXACT=1 timeout 600 python tools/perf_probe.py 512 67108864 10 slab 2:4,4:4 >> gpurun_out/c2_ab_64M.log 2>&1
timeout 300 python tools/perf_probe.py 128 1048576 20 ball 0:0,2:4,4:4 >> gpurun_out/c2_ab_64M.log 2>&1
tail -n 6 gpurun_out/c2_gpu_tests.log
cat gpurun_out/c2_ab_64M.log
