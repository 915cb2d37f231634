#!/bin/bash
# Round-end measurement on ONE B200 (about 6 GPU-minutes):  gpurun --timeout 900 -- 'bash tools/final_gpu_call.sh'
# 1. the whole GPU suite with the measured-error log  2. the default bench line (all BASELINE configs + CPU baseline)
# 3. the reference arm  4. ncu launch lists of one grounding step and one train step.   Outputs: gpurun_out/final_*
set -u
mkdir -p gpurun_out
rm -f gpurun_out/final_parity.jsonl
MPL_PARITY_LOG=gpurun_out/final_parity.jsonl python tools/gpu_pytest.py --log gpurun_out/final_tests.log tests -m gpu \
  > gpurun_out/final_tests.out 2>&1
tail -3 gpurun_out/final_tests.out
python tools/parity_table.py gpurun_out/final_parity.jsonl > gpurun_out/final_parity_errors.md 2>/dev/null
python bench.py > gpurun_out/final_all.json 2> gpurun_out/final_all.err
tail -c 300 gpurun_out/final_all.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err
tail -c 400 gpurun_out/final_reference.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/final_launches.csv python bench.py --ncu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/final_train_launches.csv python bench.py --workload train --ncu > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/final_launches.csv 2>/dev/null | head -12
python tools/launch_summary.py gpurun_out/final_train_launches.csv 2>/dev/null | head -14
