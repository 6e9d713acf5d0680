#!/bin/bash
# after the mixed-precision polar iteration in k_imp_stress: implicit tests + probe + bench line, ncu of the solve's kernels,
# and the substep capture again (profiles/traffic.json is keyed by the hash of csrc/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_implicit.py -m gpu -x -q > gpurun_out/i_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/i_tests.log
timeout 600 python tools/implicit_probe.py 128 1048576 1e-4 > gpurun_out/i_implicit_probe.log 2>&1
timeout 600 python tools/implicit_probe.py 256 4194304 1e-4 >> gpurun_out/i_implicit_probe.log 2>&1
MPM_PROBE_BASELINE=1 timeout 600 python tools/implicit_probe.py 256 4194304 1e-4 >> gpurun_out/i_implicit_probe.log 2>&1
timeout 600 python bench.py --workload implicit --steps 5 --warmup 3 > gpurun_out/i_bench_implicit.log 2>&1; echo "implicit exit $?" >> gpurun_out/i_bench_implicit.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --clock-control none --csv --metrics $M --kernel-name-base mangled -k regex:'k_imp|k_vec|k_lbfgs|k_g2p_tileILi34' -s 30 -c 120 \
  --log-file gpurun_out/i_ncu_implicit_4M.csv python tools/implicit_probe.py 256 4194304 1e-4 400 > gpurun_out/i_ncu_implicit.log 2>&1
M2=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --clock-control none --csv --metrics $M2 -s 36 -c 24 \
  --log-file gpurun_out/i_ncu_substep_64M.csv python tools/profile_step.py 512 67108864 6 > gpurun_out/i_ncu_substep.log 2>&1
tail -n 3 gpurun_out/i_tests.log; cat gpurun_out/i_implicit_probe.log | cut -c1-250; tail -n 2 gpurun_out/i_bench_implicit.log | cut -c1-300
