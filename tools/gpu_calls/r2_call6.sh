#!/bin/bash
mkdir -p gpurun_out
L=realtime-deformations_b200
MPM_B200_LIB=$L/libmpm_b200_prof.so timeout 300 python tools/p2g_phase_profile.py run 512 67108864 2:4 >> gpurun_out/c6_phase.log 2>&1
timeout 600 python tools/perf_probe.py 512 67108864 10 slab 2:4,4:4 > gpurun_out/c6_ab_64M.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c6_gpu_tests.log 2>&1
cat gpurun_out/c6_phase.log gpurun_out/c6_ab_64M.log; tail -3 gpurun_out/c6_gpu_tests.log
