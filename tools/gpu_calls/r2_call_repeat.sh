#!/bin/bash
# repeatability: implicit tests five times, whole GPU suite twice
mkdir -p gpurun_out
for i in 1 2 3 4 5; do timeout 300 python -m pytest tests/test_gpu_implicit.py -m gpu -q > gpurun_out/t_imp_$i.log 2>&1; echo "implicit run $i exit $?"; tail -n 1 gpurun_out/t_imp_$i.log; done
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all_$i.log 2>&1; echo "suite run $i exit $?"; tail -n 1 gpurun_out/t_all_$i.log; done
