#!/bin/bash
# the A/B kernel pairings (F-update as its own kernel, tile / baseline mixes) through the whole GPU suite on the final library
mkdir -p gpurun_out
MPM_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_exp_tests.log 2>&1; echo "experimental-variant tests exit $?" >> gpurun_out/r_exp_tests.log
tail -n 4 gpurun_out/r_exp_tests.log
