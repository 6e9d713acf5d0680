#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/c7_bench_c5.log 2>&1; echo "c5 exit $?" >> gpurun_out/c7_bench_c5.log
for c in 1 2 3 4; do
  timeout 300 python bench.py --config $c --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/c7_bench_c$c.log 2>&1; echo "c$c exit $?" >> gpurun_out/c7_bench_c$c.log
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c7_gpu_tests.log 2>&1
for c in 5 1 2 3 4; do tail -n 2 gpurun_out/c7_bench_c$c.log | cut -c1-700; done
tail -5 gpurun_out/c7_gpu_tests.log
