#!/bin/bash
# scaling run on one 8-GPU box: N = 8, 4, 2 (strong scaling of config 5), each with the in-run multi-vs-single cross-check
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/g8_ngpu.txt
for n in 8 4 2; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) \
    bench.py --gpus $n --steps 50 --warmup 10 > gpurun_out/g8_bench_n$n.log 2>&1
  echo "n=$n exit $?" >> gpurun_out/g8_bench_n$n.log
done
MPM_B200_PEER_HALO=0 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29820 \
    bench.py --gpus 8 --steps 50 --warmup 10 --no-multi-check > gpurun_out/g8_bench_n8_nccl.log 2>&1
echo "n=8 nccl exit $?" >> gpurun_out/g8_bench_n8_nccl.log
for f in gpurun_out/g8_bench_n*.log; do echo == $f; tail -n 2 $f | cut -c1-400; done
