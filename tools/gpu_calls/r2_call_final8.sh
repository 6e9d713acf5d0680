#!/bin/bash
# bench lines of all configs, reference arm, implicit line, smoke on the final sources
mkdir -p gpurun_out
timeout 400 python bench.py --steps 50 --warmup 10 > gpurun_out/z_bench_c5.log 2>&1; echo "c5 exit $?" >> gpurun_out/z_bench_c5.log
for c in 1 2 3 4; do timeout 400 python bench.py --config $c --steps 50 --warmup 10 > gpurun_out/z_bench_c$c.log 2>&1; echo "c$c exit $?" >> gpurun_out/z_bench_c$c.log; done
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/z_bench_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/z_bench_ref.log
timeout 600 python bench.py --workload implicit --steps 5 --warmup 3 > gpurun_out/z_bench_implicit.log 2>&1; echo "implicit exit $?" >> gpurun_out/z_bench_implicit.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/z_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/z_smoke.log
for c in 5 1 2 3 4; do tail -n 2 gpurun_out/z_bench_c$c.log | cut -c1-230; done; tail -n 1 gpurun_out/z_bench_ref.log; tail -n 2 gpurun_out/z_bench_implicit.log | cut -c1-200; tail -n 2 gpurun_out/z_smoke.log
