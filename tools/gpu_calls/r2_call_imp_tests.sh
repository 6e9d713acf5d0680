#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_implicit.py -m gpu -x -q --durations=4 > gpurun_out/n_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/n_tests.log
tail -n 9 gpurun_out/n_tests.log
