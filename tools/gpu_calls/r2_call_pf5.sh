#!/bin/bash
# A/B: L2 prefetch of the next grid block in k_grid_update
mkdir -p gpurun_out; rm -f gpurun_out/a_probe.log
for v in base gupf base gupf; do
  echo "== $v" >> gpurun_out/a_probe.log
  MPM_B200_LIB=$PWD/realtime-deformations_b200/libmpm_b200_$v.so timeout 300 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/a_probe.log 2>&1
done
cat gpurun_out/a_probe.log | cut -c90-220
