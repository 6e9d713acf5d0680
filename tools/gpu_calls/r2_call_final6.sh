#!/bin/bash
# headline line and implicit line on the final sources with the matching traffic.json
mkdir -p gpurun_out
timeout 400 python bench.py --steps 50 --warmup 10 > gpurun_out/m_bench_c5.log 2>&1; echo "c5 exit $?" >> gpurun_out/m_bench_c5.log
timeout 600 python bench.py --workload implicit --steps 5 --warmup 3 > gpurun_out/m_bench_implicit.log 2>&1; echo "implicit exit $?" >> gpurun_out/m_bench_implicit.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/m_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/m_smoke.log
tail -n 2 gpurun_out/m_bench_c5.log | cut -c1-300; tail -n 2 gpurun_out/m_bench_implicit.log | cut -c1-200; tail -n 2 gpurun_out/m_smoke.log
