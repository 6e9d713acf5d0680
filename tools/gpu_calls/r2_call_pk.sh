#!/bin/bash
# A/B of the packed two-particle tolerance-form F-update inside P2G (MPM_B200_FUPD_PACKED=1)
mkdir -p gpurun_out
MPM_B200_FUPD_PACKED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trajectory or staged_and_fused or synthetic_ball or config2_family or ragged or material_sweep" > gpurun_out/pk_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/pk_tests.log
for f in 0 1 0 1; do
MPM_B200_FUPD_PACKED=$f timeout 300 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/pk_probe.log 2>&1
done
for f in 0 1; do
MPM_B200_FUPD_PACKED=$f timeout 300 python tools/perf_probe.py 256 8388608 20 ball 0:0 >> gpurun_out/pk_probe.log 2>&1
done
tail -n 3 gpurun_out/pk_tests.log; cat gpurun_out/pk_probe.log | cut -c1-260
