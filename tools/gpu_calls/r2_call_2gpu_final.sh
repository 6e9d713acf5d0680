#!/bin/bash
# the multi-GPU path once more on the final library: 2-GPU test of the suite, multi-vs-single-domain check, bench lines (config 5 and 3)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or peer_memory" > gpurun_out/j_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/j_tests.log
timeout 300 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/multi_check.py 128 2097152 30 > gpurun_out/j_multi_check.log 2>&1
echo "multi_check exit $?" >> gpurun_out/j_multi_check.log
timeout 420 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/j_bench_c5.log 2>&1
echo "bench c5 exit $?" >> gpurun_out/j_bench_c5.log
timeout 420 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --config 3 --steps 50 --warmup 10 > gpurun_out/j_bench_c3.log 2>&1
echo "bench c3 exit $?" >> gpurun_out/j_bench_c3.log
tail -n 3 gpurun_out/j_tests.log; tail -n 4 gpurun_out/j_multi_check.log | cut -c1-300
for f in gpurun_out/j_bench_*.log; do echo == $f; tail -n 2 $f | cut -c1-400; done
