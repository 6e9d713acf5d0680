#!/bin/bash
# after profiles/traffic.json was regenerated for the frozen sources: the headline line once more (now carrying `traffic`), and the
# implicit solve's kernels under ncu (mangled names, so that the gradient-gather instantiation of k_g2p_tile can be told apart)
mkdir -p gpurun_out
timeout 400 python bench.py --steps 50 --warmup 10 > gpurun_out/h_bench_c5.log 2>&1; echo "c5 exit $?" >> gpurun_out/h_bench_c5.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --clock-control none --csv --metrics $M --kernel-name-base mangled -k regex:'k_imp|k_vec|k_lbfgs|k_g2p_tileILi34' -s 30 -c 120 \
  --log-file gpurun_out/h_ncu_implicit_4M.csv python tools/implicit_probe.py 256 4194304 1e-4 400 > gpurun_out/h_ncu_implicit.log 2>&1
tail -n 2 gpurun_out/h_bench_c5.log | cut -c1-300; tail -n 3 gpurun_out/h_ncu_implicit.log | cut -c1-200; wc -l gpurun_out/h_ncu_implicit_4M.csv
