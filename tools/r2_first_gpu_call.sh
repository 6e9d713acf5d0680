#!/bin/bash
# First GPU call of the next round: confirm what was changed or prepared after round 1's GPU budget ran out.
#   gpurun --timeout 1800 -- 'bash tools/r2_first_gpu_call.sh'
# Everything is wrapped in `timeout`; outputs land in gpurun_out/.
mkdir -p gpurun_out
# 1. parity suite: default path (rotated P2G record walk, single-quotient weights, block coordinates in the work items)
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/gpu_tests.log 2>&1
echo "default gpu tests: exit $?" | tee -a gpurun_out/gpu_tests.log
# 2. A/B of the rotated record walk at 64 Mi (third field 0 = the aligned walk measured in round 1); one process, one scene
timeout 420 python tools/perf_probe.py 512 67108864 10 slab 0:0:0,0:0:1 > gpurun_out/rotate_64M.log 2>&1
# 3. shared-memory wavefronts / bank conflicts of P2G with and without the rotation (model: tests/emu smem profile)
for rot in 0 1; do
  MPM_B200_P2G_ROTATE=$rot timeout 400 ncu --clock-control none -k regex:k_p2g_tile -c 2 --csv \
    --metrics gpu__time_duration.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --log-file gpurun_out/ncu_p2g_rotate_${rot}.csv python tools/profile_step.py 512 67108864 2 > gpurun_out/ncu_p2g_rotate_${rot}.log 2>&1
done
# 3b. one full ncu capture of the three top kernels of the default path (source-level counters: compile has -lineinfo)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_p2g_tile|k_g2p_tile|k_fupdate' -s 4 -c 3 \
  -o gpurun_out/ncu_r2_top_kernels_64M python tools/profile_step.py 512 67108864 3 > gpurun_out/ncu_full.log 2>&1
# 4. experimental kernels: parity, then A/B timing
timeout 900 env MPM_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/exp_tests.log 2>&1
echo "experimental gpu tests: exit $?" | tee -a gpurun_out/exp_tests.log
timeout 300 python tools/perf_probe.py 256 8388608 20 slab 0:0,0:2,0:3,0:4,2:0,3:0,4:4 > gpurun_out/ab_8M.log 2>&1
timeout 480 python tools/perf_probe.py 512 67108864 10 slab 0:0,4:4,3:0,2:0,0:4 > gpurun_out/ab_64M.log 2>&1
# 4b. compute-sanitizer on the small case: default kernels (changed after round 1's sanitizer run) and all options on
for v in "0 0" "4 4"; do
  for tool in memcheck racecheck; do
    echo "== $tool, variants $v" >> gpurun_out/sanitizer_r2.txt
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_case.py $v 2>&1 | grep -v "^=========     \|^=========$" | tail -8 >> gpurun_out/sanitizer_r2.txt
  done
done
# 5. (needs `gpurun --gpus 2`) peer-memory halo: the single-process protocol test is part of step 4; then 2 ranks with
#    CUDA IPC against one domain, with and without MPM_B200_PEER_HALO, every multi-rank command under `timeout`
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for peer in 0 1; do
    MPM_B200_PEER_HALO=$peer timeout 300 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
      tools/multi_check.py 128 2097152 30 > gpurun_out/multi_check_peer_${peer}.log 2>&1
    MPM_B200_PEER_HALO=$peer timeout 420 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
      bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_peer_${peer}.log 2>&1
  done
  tail -n 3 gpurun_out/multi_check_peer_*.log gpurun_out/bench_2gpu_peer_*.log
fi
tail -n 8 gpurun_out/gpu_tests.log gpurun_out/rotate_64M.log gpurun_out/exp_tests.log gpurun_out/ab_8M.log gpurun_out/ab_64M.log
