#!/bin/bash
# e2e at N = 8: NUMA-local pinned buffers (bench.py binds each rank to its GPU's CPUs) against no binding
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/g8c_topo.txt 2>&1
for b in 0 1; do
  MPM_B200_NUMA_BIND=$b timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29830 + b)) \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-multi-check > gpurun_out/g8c_bench_bind$b.log 2>&1
  echo "bind=$b exit $?" >> gpurun_out/g8c_bench_bind$b.log
done
head -30 gpurun_out/g8c_topo.txt | cut -c1-200
for f in gpurun_out/g8c_bench_bind*.log; do python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e'].get('host_numa_binding'))
PY
done
