#!/bin/bash
# scaling run on one 8-GPU box: N = 8, 4, 2, 1 (strong scaling of config 5), each with the in-run checks
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) \
    bench.py --gpus $n --steps 50 --warmup 10 > gpurun_out/g8b_bench_n$n.log 2>&1
  echo "n=$n exit $?" >> gpurun_out/g8b_bench_n$n.log
done
timeout 420 python bench.py --gpus 1 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/g8b_bench_n1.log 2>&1
echo "n=1 exit $?" >> gpurun_out/g8b_bench_n1.log
for f in gpurun_out/g8b_bench_n*.log; do echo == $f; tail -n 2 $f | cut -c1-330; done
