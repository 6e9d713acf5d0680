#!/bin/bash
# final evidence of the round on the frozen kernels: bench lines of all five configs, ncu per-kernel metrics over whole substeps
mkdir -p gpurun_out
timeout 400 python bench.py --steps 50 --warmup 10 > gpurun_out/f_bench_c5.log 2>&1; echo "c5 exit $?" >> gpurun_out/f_bench_c5.log
for c in 1 2 3 4; do timeout 400 python bench.py --config $c --steps 50 --warmup 10 > gpurun_out/f_bench_c$c.log 2>&1; echo "c$c exit $?" >> gpurun_out/f_bench_c$c.log; done
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/f_bench_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/f_bench_ref.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --clock-control none --csv --metrics $M -s 36 -c 24 \
  --log-file gpurun_out/f_ncu_substep_64M.csv python tools/profile_step.py 512 67108864 6 > gpurun_out/f_ncu_substep.log 2>&1
for c in 5 1 2 3 4 ref; do tail -n 2 gpurun_out/f_bench_$c.log | cut -c1-260; done
