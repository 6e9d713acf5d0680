#!/bin/bash
# first hardware run of the single-pass substep: parity subset, then A/B at 64 Mi against the two-kernel form
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trajectory or staged_and_fused or parked or ragged or synthetic_ball or config2_family or invariants" > gpurun_out/sp1_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/sp1_tests.log
timeout 300 python tools/perf_probe.py 512 67108864 10 slab 0:0 > gpurun_out/sp1_probe.log 2>&1
MPM_PROBE_SUBSTEP_FORM=1 timeout 300 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/sp1_probe.log 2>&1
MPM_PROBE_SHUFFLE=1 timeout 300 python tools/perf_probe.py 256 8388608 10 slab 0:0 >> gpurun_out/sp1_probe.log 2>&1
MPM_PROBE_SHUFFLE=1 MPM_PROBE_SUBSTEP_FORM=1 timeout 300 python tools/perf_probe.py 256 8388608 10 slab 0:0 >> gpurun_out/sp1_probe.log 2>&1
timeout 300 python tools/perf_probe.py 256 8388608 20 ball 0:0 >> gpurun_out/sp1_probe.log 2>&1
MPM_PROBE_SUBSTEP_FORM=1 timeout 300 python tools/perf_probe.py 256 8388608 20 ball 0:0 >> gpurun_out/sp1_probe.log 2>&1
tail -n 4 gpurun_out/sp1_tests.log; cat gpurun_out/sp1_probe.log | cut -c1-330
