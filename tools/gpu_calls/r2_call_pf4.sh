#!/bin/bash
# A/B: streaming (evict-first) stores of the planes written for the next substep (st), streaming loads of the F-update inputs (ld)
mkdir -p gpurun_out; rm -f gpurun_out/x_probe.log
for v in base st ld stld base st ld stld; do
  echo "== $v" >> gpurun_out/x_probe.log
  MPM_B200_LIB=$PWD/realtime-deformations_b200/libmpm_b200_$v.so timeout 300 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/x_probe.log 2>&1
done
cat gpurun_out/x_probe.log | cut -c90-220
