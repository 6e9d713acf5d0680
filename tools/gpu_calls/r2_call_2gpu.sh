#!/bin/bash
# 2-GPU validation of the slab decomposition: message halo (NCCL send/recv) vs ghost-layer reduction over peer memory
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g2_smi.txt
for peer in 0 1; do
  MPM_B200_PEER_HALO=$peer timeout 300 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    tools/multi_check.py 128 2097152 30 > gpurun_out/g2_multi_check_peer_${peer}.log 2>&1
  echo "multi_check peer=$peer exit $?" >> gpurun_out/g2_multi_check_peer_${peer}.log
  MPM_B200_PEER_HALO=$peer timeout 420 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/g2_bench_peer_${peer}.log 2>&1
  echo "bench peer=$peer exit $?" >> gpurun_out/g2_bench_peer_${peer}.log
done
timeout 300 python -m pytest tests -m gpu -q -k "two_gpus or slab" > gpurun_out/g2_tests.log 2>&1
tail -n 6 gpurun_out/g2_multi_check_peer_*.log
for f in gpurun_out/g2_bench_peer_*.log; do echo == $f; tail -n 3 $f | cut -c1-1500; done
tail -n 4 gpurun_out/g2_tests.log
