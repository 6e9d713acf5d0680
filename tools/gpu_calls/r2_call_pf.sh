#!/bin/bash
# A/B: L2 prefetch of the F-update planes in P2G's derive phase (pff), of the next block's P2G planes after the fold (pfn), both
mkdir -p gpurun_out; rm -f gpurun_out/u_probe.log
for v in base pff pfn both base both; do
  echo "== $v" >> gpurun_out/u_probe.log
  MPM_B200_LIB=$PWD/realtime-deformations_b200/libmpm_b200_$v.so timeout 300 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/u_probe.log 2>&1
done
for v in base both; do
  echo "== $v (ball 8 Mi)" >> gpurun_out/u_probe.log
  MPM_B200_LIB=$PWD/realtime-deformations_b200/libmpm_b200_$v.so timeout 300 python tools/perf_probe.py 256 8388608 20 ball 0:0 >> gpurun_out/u_probe.log 2>&1
done
cat gpurun_out/u_probe.log | cut -c1-250
