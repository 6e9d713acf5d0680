#!/bin/bash
# compute-sanitizer on the final library: fused + staged substeps and an implicit solve that really iterates
mkdir -p gpurun_out; rm -f gpurun_out/k_sanitizer.txt
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool, default kernels incl. a 12-iteration implicit solve (tools/sanitize_case.py 0 0)" >> gpurun_out/k_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_case.py 0 0 2>&1 | grep -v "^=========     \|^=========$" | tail -8 >> gpurun_out/k_sanitizer.txt
done
cat gpurun_out/k_sanitizer.txt
