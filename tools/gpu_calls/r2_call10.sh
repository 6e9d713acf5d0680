#!/bin/bash
# parity at size + sanitizer + fresh ncu evidence of the current default kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/c10_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/c10_gpu_tests.log
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool, default kernels (tools/sanitize_case.py 0 0)" >> gpurun_out/c10_sanitizer.txt
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_case.py 0 0 2>&1 | grep -v "^=========     \|^=========$" | tail -8 >> gpurun_out/c10_sanitizer.txt
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --clock-control none --csv --metrics $M -s 40 -c 26 \
  --log-file gpurun_out/c10_ncu_substep_64M.csv python tools/profile_step.py 512 67108864 6 > gpurun_out/c10_ncu_substep.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_p2g_tile|k_g2p_tile' -s 6 -c 2 \
  -o gpurun_out/c10_ncu_full_64M python tools/profile_step.py 512 67108864 5 > gpurun_out/c10_ncu_full.log 2>&1
tail -n 14 gpurun_out/c10_gpu_tests.log; cat gpurun_out/c10_sanitizer.txt; ls -la gpurun_out | tail -5
