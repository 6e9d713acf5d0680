#!/bin/bash
# implicit time integration on hardware: parity tests, then timing at 1 Mi (config 2 scene) and 4 Mi particles
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_implicit.py -m gpu -x -q > gpurun_out/imp1_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/imp1_tests.log
timeout 600 python tools/implicit_probe.py 128 1048576 1e-4 > gpurun_out/imp1_probe.log 2>&1
timeout 600 python tools/implicit_probe.py 256 4194304 1e-4 >> gpurun_out/imp1_probe.log 2>&1
MPM_PROBE_BASELINE=1 timeout 600 python tools/implicit_probe.py 256 4194304 1e-4 >> gpurun_out/imp1_probe.log 2>&1
tail -n 4 gpurun_out/imp1_tests.log; cat gpurun_out/imp1_probe.log | cut -c1-300
