#!/bin/bash
# ncu --set full capture of one launch of the single-pass kernel (8 Mi particles / 256^3 slab)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'^k_g2p2g$' -s 3 -c 1 \
  -o gpurun_out/sp2_ncu_full python tools/profile_step.py 256 8388608 5 > gpurun_out/sp2_ncu_full.log 2>&1
tail -n 3 gpurun_out/sp2_ncu_full.log
