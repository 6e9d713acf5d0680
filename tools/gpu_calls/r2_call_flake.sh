#!/bin/bash
# how close do the calibrated trajectory tests run to their bounds? five repetitions, error / noise-floor ratios collected
mkdir -p gpurun_out; rm -f gpurun_out/s_margins.log
for i in 1 2 3 4 5; do
  MPM_TEST_MARGIN_LOG=gpurun_out/s_margins.log timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "synthetic_ball or quadratic or material_sweep or sphere_collider or config2_family or eighth or config2_full_size_vs or rotated or deterministic" > gpurun_out/s_tests_$i.log 2>&1
  echo "run $i exit $?" >> gpurun_out/s_tests_$i.log; tail -n 2 gpurun_out/s_tests_$i.log
done
wc -l gpurun_out/s_margins.log
