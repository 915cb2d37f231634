#!/bin/bash
# Decode-kernel experiment: parity tests of everything that runs the persistent decode kernel, then per-phase timings
# and step times at B = 1 / 8 for a list of L2 look-ahead settings.  Usage: bash tools/dk_sweep.sh "LA:SPEC:EVICT:LA2 ..."
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stacks_gpu.py tests/test_model_gpu.py tests/test_zz_serve_gpu.py -m gpu -q -x \
  -p no:cacheprovider --tb=short > gpurun_out/dk_tests.log 2>&1
tail -4 gpurun_out/dk_tests.log
for cfg in ${1:-"0:4:0 16:4:0"}; do
  IFS=: read la spec evict la2 <<< "$cfg"
  for B in ${DK_BS:-1 8}; do
    MPL_DK_LA=$la MPL_DK_SPEC=$spec MPL_DK_EVICT=$evict MPL_DK_LA2=${la2:-16} timeout 150 python tests/dev/dev_llama.py timing $B \
      > gpurun_out/dk_${la}_${spec}_${evict}_${la2:-16}_B$B.log 2>&1
    grep -E "ms/step|total" gpurun_out/dk_${la}_${spec}_${evict}_${la2:-16}_B$B.log
  done
done
