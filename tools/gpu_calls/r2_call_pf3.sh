#!/bin/bash
# A/B: P2G prefetch of the next CHUNK of a block (chunk), gather with its work item one block ahead + L2 prefetch (g2p), both (all)
mkdir -p gpurun_out; rm -f gpurun_out/w_probe.log
for v in both chunk g2p all both chunk g2p all; do
  echo "== $v" >> gpurun_out/w_probe.log
  MPM_B200_LIB=$PWD/realtime-deformations_b200/libmpm_b200_$v.so timeout 300 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/w_probe.log 2>&1
done
cat gpurun_out/w_probe.log | cut -c90-220
