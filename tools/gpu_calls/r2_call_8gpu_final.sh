#!/bin/bash
# scaling points N = 8 and N = 4 on the final library (one 8-GPU box)
mkdir -p gpurun_out
for n in 8 4; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29840 + n)) \
    bench.py --gpus $n --steps 50 --warmup 10 > gpurun_out/p_bench_${n}gpu.log 2>&1
  echo "n=$n exit $?" >> gpurun_out/p_bench_${n}gpu.log
done
for n in 8 4; do tail -n 2 gpurun_out/p_bench_${n}gpu.log | cut -c1-330; done
