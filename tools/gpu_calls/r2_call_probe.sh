#!/bin/bash
# quick A/B probe of the current build at 64 Mi and 8 Mi (tools/perf_probe.py), plus the fused-path parity subset
mkdir -p gpurun_out; rm -f gpurun_out/q_probe.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trajectory or staged_and_fused or synthetic_ball or ragged" > gpurun_out/q_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/q_tests.log
for i in 1 2; do timeout 300 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/q_probe.log 2>&1; done
timeout 300 python tools/perf_probe.py 256 8388608 20 ball 0:0 >> gpurun_out/q_probe.log 2>&1
tail -n 3 gpurun_out/q_tests.log; cat gpurun_out/q_probe.log | cut -c1-260
