#!/bin/bash
# final sources of the round (P2G with L2 prefetches): GPU suite, sanitizer, substep capture for traffic.json
mkdir -p gpurun_out; rm -f gpurun_out/y_sanitizer.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/y_gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/y_gpu_tests.log
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool, default kernels incl. a 12-iteration implicit solve (tools/sanitize_case.py 0 0)" >> gpurun_out/y_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_case.py 0 0 2>&1 | grep -v "^=========     \|^=========$" | tail -8 >> gpurun_out/y_sanitizer.txt
done
M2=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --clock-control none --csv --metrics $M2 -s 36 -c 24 \
  --log-file gpurun_out/y_ncu_substep_64M.csv python tools/profile_step.py 512 67108864 6 > gpurun_out/y_ncu_substep.log 2>&1
tail -n 3 gpurun_out/y_gpu_tests.log; cat gpurun_out/y_sanitizer.txt; tail -n 1 gpurun_out/y_ncu_substep.log
