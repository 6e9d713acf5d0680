#!/bin/bash
# A/B: rank-addressed early prefetch of the next block's planes (rank), rank + id-addressed (rankboth), against id-addressed (both)
mkdir -p gpurun_out; rm -f gpurun_out/v_probe.log
for v in both rank rankboth both rank rankboth; do
  echo "== $v" >> gpurun_out/v_probe.log
  MPM_B200_LIB=$PWD/realtime-deformations_b200/libmpm_b200_$v.so timeout 300 python tools/perf_probe.py 512 67108864 10 slab 0:0 >> gpurun_out/v_probe.log 2>&1
done
for v in both rank; do
  echo "== $v (ball 8 Mi, shuffled upload)" >> gpurun_out/v_probe.log
  MPM_PROBE_SHUFFLE=1 MPM_B200_LIB=$PWD/realtime-deformations_b200/libmpm_b200_$v.so timeout 300 python tools/perf_probe.py 256 8388608 20 ball 0:0 >> gpurun_out/v_probe.log 2>&1
done
cat gpurun_out/v_probe.log | cut -c90-220
