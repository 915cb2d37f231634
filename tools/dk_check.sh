#!/bin/bash
# Decode-kernel change: parity of everything that runs the persistent decode kernel, then per-phase timings at B = 8 / 1.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_stacks_gpu.py tests/test_model_gpu.py tests/test_zz_serve_gpu.py tests/test_fullwidth_gpu.py -m gpu -q -x \
  -p no:cacheprovider --tb=short > gpurun_out/dk_tests.log 2>&1
tail -3 gpurun_out/dk_tests.log
for B in 8 1; do timeout 200 python tests/dev/dev_llama.py timing $B > gpurun_out/dk_now_B$B.log 2>&1; head -${DK_LINES:-22} gpurun_out/dk_now_B$B.log; done
