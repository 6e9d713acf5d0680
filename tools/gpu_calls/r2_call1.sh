#!/bin/bash
# Round 2, GPU call 1: measure what exists (trimmed from r2_first_gpu_call.sh to fit ~20 box-minutes).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem --format=csv > gpurun_out/c1_smi.txt 2>&1
# 1. default parity suite
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/c1_gpu_tests.log 2>&1
echo "default gpu tests: exit $?" >> gpurun_out/c1_gpu_tests.log
# 2. A/B at 64 Mi: every unmeasured variant, one process (scene cached)
timeout 600 python tools/perf_probe.py 512 67108864 10 slab 0:0:0,0:0,2:0,3:0,0:2,0:3,0:4,4:4 > gpurun_out/c1_ab_64M.log 2>&1
# 3. ncu: per-kernel metric set over one whole substep (all kernels), 64 Mi
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_mio_throttle.ratio,smsp__average_warp_latency_issue_stalled_lg_throttle.ratio,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,smsp__average_warp_latency_issue_stalled_not_selected.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__average_warp_latency_issue_stalled_branch_resolving.ratio,smsp__average_warp_latency_issue_stalled_membar.ratio
timeout 600 ncu --clock-control none --csv --metrics $M -s 40 -c 16 \
  --log-file gpurun_out/c1_ncu_substep_64M.csv python tools/profile_step.py 512 67108864 5 > gpurun_out/c1_ncu_substep.log 2>&1
# 4. one full capture of the three top kernels (source-level)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_p2g_tile|k_g2p_tile|k_fupdate' -s 6 -c 3 \
  -o gpurun_out/c1_ncu_full_64M python tools/profile_step.py 512 67108864 4 > gpurun_out/c1_ncu_full.log 2>&1
# 5. experimental variants: parity
timeout 600 env MPM_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/c1_exp_tests.log 2>&1
echo "experimental gpu tests: exit $?" >> gpurun_out/c1_exp_tests.log
# 6. bench default (contract line)
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/c1_bench.log 2>&1
tail -n 4 gpurun_out/c1_gpu_tests.log gpurun_out/c1_exp_tests.log
cat gpurun_out/c1_ab_64M.log
tail -n 2 gpurun_out/c1_bench.log
ls -la gpurun_out
