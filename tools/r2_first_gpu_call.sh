#!/bin/bash
# First GPU call of the next round: validate the experimental kernels prepared at the end of round 1 and A/B them.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_call.sh'
# Everything is wrapped in `timeout`; outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 600 env MPM_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu -x -q > gpurun_out/exp_tests.log 2>&1
echo "experimental gpu tests: exit $?" | tee -a gpurun_out/exp_tests.log
timeout 300 python tools/perf_probe.py 256 8388608 20 slab 0:0,0:2,0:3,0:4,2:0,3:0,4:4 > gpurun_out/ab_8M.log 2>&1
timeout 480 python tools/perf_probe.py 512 67108864 10 slab 0:0,4:4,3:0,2:0,0:4 > gpurun_out/ab_64M.log 2>&1
tail -n 8 gpurun_out/exp_tests.log gpurun_out/ab_8M.log gpurun_out/ab_64M.log
