#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/g2b_tests.log 2>&1
for peer in 1 0; do
  MPM_B200_PEER_HALO=$peer timeout 300 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    tools/multi_check.py 128 2097152 30 > gpurun_out/g2b_multi_check_peer_${peer}.log 2>&1
  echo "multi_check peer=$peer exit $?" >> gpurun_out/g2b_multi_check_peer_${peer}.log
  MPM_B200_PEER_HALO=$peer timeout 420 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/g2b_bench_peer_${peer}.log 2>&1
  echo "bench peer=$peer exit $?" >> gpurun_out/g2b_bench_peer_${peer}.log
done
timeout 420 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 \
    bench.py --gpus 2 --config 3 --steps 50 --warmup 10 > gpurun_out/g2b_bench_c3.log 2>&1
echo "bench c3 exit $?" >> gpurun_out/g2b_bench_c3.log
tail -n 4 gpurun_out/g2b_tests.log
tail -n 5 gpurun_out/g2b_multi_check_peer_*.log
for f in gpurun_out/g2b_bench_*.log; do echo == $f; tail -n 2 $f | cut -c1-300; done
