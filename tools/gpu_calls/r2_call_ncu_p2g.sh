#!/bin/bash
# ncu --set full of the final P2G + F-update kernel (with the L2 prefetches), one launch at 64 Mi
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_p2g_tile' -s 3 -c 1 \
  -o gpurun_out/b_ncu_p2g_final python tools/profile_step.py 512 67108864 5 > gpurun_out/b_ncu_p2g.log 2>&1
tail -n 2 gpurun_out/b_ncu_p2g.log; ls -la gpurun_out/b_ncu_p2g_final.ncu-rep
