#!/bin/bash
# f-4 evidence on ONE B200: the geo parity tests + the rank-16 stage-2 train test, timing of the full-size sampler, an
# `ncu --set full` capture of its kernels (read here with tools/ncu_summary.py).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_geo_gpu.py tests/test_train_gpu.py -m gpu -q -p no:cacheprovider --tb=short \
  -k "geo or rank_16 or stage2" > gpurun_out/geo_tests.log 2>&1
tail -25 gpurun_out/geo_tests.log
timeout 120 python tools/geo_run.py --regions 4 > gpurun_out/geo_time_r4.json 2> gpurun_out/geo_time.err
timeout 120 python tools/geo_run.py --regions 16 > gpurun_out/geo_time_r16.json 2>> gpurun_out/geo_time.err
cat gpurun_out/geo_time_r4.json gpurun_out/geo_time_r16.json
timeout 400 ncu --set full --clock-control none -k regex:geo_ -c 14 -f -o gpurun_out/r02_geo \
  python tools/geo_run.py --regions 4 --once > gpurun_out/geo_ncu.log 2>&1
tail -2 gpurun_out/geo_ncu.log
