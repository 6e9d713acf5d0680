#!/bin/bash
# final evidence of round 2 on the frozen sources: full GPU suite, sanitizer, bench lines of all five configs + the reference arm,
# ncu per-kernel metrics over whole substeps (-> profiles/traffic.json), implicit-integration probe and its kernels under ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/g_gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/g_gpu_tests.log
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool, default kernels incl. implicit time integration (tools/sanitize_case.py 0 0)" >> gpurun_out/g_sanitizer.txt
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_case.py 0 0 2>&1 | grep -v "^=========     \|^=========$" | tail -8 >> gpurun_out/g_sanitizer.txt
done
timeout 400 python bench.py --steps 50 --warmup 10 > gpurun_out/g_bench_c5.log 2>&1; echo "c5 exit $?" >> gpurun_out/g_bench_c5.log
for c in 1 2 3 4; do timeout 400 python bench.py --config $c --steps 50 --warmup 10 > gpurun_out/g_bench_c$c.log 2>&1; echo "c$c exit $?" >> gpurun_out/g_bench_c$c.log; done
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/g_bench_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/g_bench_ref.log
timeout 600 python bench.py --workload implicit --steps 5 --warmup 3 > gpurun_out/g_bench_implicit.log 2>&1; echo "implicit exit $?" >> gpurun_out/g_bench_implicit.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --clock-control none --csv --metrics $M -s 36 -c 24 \
  --log-file gpurun_out/g_ncu_substep_64M.csv python tools/profile_step.py 512 67108864 6 > gpurun_out/g_ncu_substep.log 2>&1
timeout 600 python tools/implicit_probe.py 128 1048576 1e-4 > gpurun_out/g_implicit_probe.log 2>&1
timeout 600 python tools/implicit_probe.py 256 4194304 1e-4 >> gpurun_out/g_implicit_probe.log 2>&1
MPM_PROBE_BASELINE=1 timeout 600 python tools/implicit_probe.py 256 4194304 1e-4 >> gpurun_out/g_implicit_probe.log 2>&1
timeout 600 ncu --clock-control none --csv --metrics $M -k regex:'k_imp|k_g2p_tile|k_vec|k_lbfgs' -s 40 -c 60 \
  --log-file gpurun_out/g_ncu_implicit_4M.csv python tools/implicit_probe.py 256 4194304 1e-4 400 > gpurun_out/g_ncu_implicit.log 2>&1
tail -n 10 gpurun_out/g_gpu_tests.log; cat gpurun_out/g_sanitizer.txt
for c in 5 1 2 3 4 ref; do tail -n 2 gpurun_out/g_bench_c$c.log 2>/dev/null | cut -c1-260; done; tail -n 2 gpurun_out/g_bench_ref.log | cut -c1-260
cat gpurun_out/g_implicit_probe.log | cut -c1-260; tail -n 2 gpurun_out/g_bench_implicit.log | cut -c1-400
