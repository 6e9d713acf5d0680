#!/bin/bash
mkdir -p gpurun_out
# value of P2G's cell-ordering rewrite of sorted_ids: steady state on the generated (cell-major) order and on a shuffled upload
for sh in 0 1; do for ro in 1 0; do
  echo "== shuffle=$sh reorder_ids=$ro" >> gpurun_out/c8_reorder.log
  MPM_PROBE_SHUFFLE=$sh MPM_B200_P2G_REORDER_IDS=$ro timeout 600 python tools/perf_probe.py 256 8388608 40 slab 0:0 >> gpurun_out/c8_reorder.log 2>&1
done; done
cat gpurun_out/c8_reorder.log
