#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "quadratic or stage_level or synthetic_ball or default_scene or fupdate" > gpurun_out/c15_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/c15_gpu_tests.log
timeout 600 python tools/perf_probe.py 512 67108864 10 slab 0:0 > gpurun_out/c15_perf.log 2>&1
MPM_PROBE_STENCIL=1 timeout 600 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/c15_perf.log 2>&1
tail -n 8 gpurun_out/c15_gpu_tests.log; cat gpurun_out/c15_perf.log
