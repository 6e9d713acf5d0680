#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/c14_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/c14_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c14_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/c14_smoke.log
for c in 1 2; do timeout 400 python bench.py --config $c --steps 50 --warmup 10 > gpurun_out/c14_bench_c$c.log 2>&1; echo "c$c exit $?" >> gpurun_out/c14_bench_c$c.log; done
tail -n 12 gpurun_out/c14_gpu_tests.log; tail -3 gpurun_out/c14_smoke.log
for c in 1 2; do tail -n 2 gpurun_out/c14_bench_c$c.log | cut -c1-200; done
