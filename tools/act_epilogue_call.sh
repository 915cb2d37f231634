#!/bin/bash
# Activation-epilogue change: GEMM / stack / model / train parity, grounding + ICL + train + decode bench lines, launch list.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_stacks_gpu.py tests/test_model_gpu.py tests/test_mask_train_gpu.py \
  tests/test_fullwidth_gpu.py tests/test_train_gpu.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/act_tests.log 2>&1
tail -4 gpurun_out/act_tests.log
for w in grounding icl train; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/act_bench_$w.json 2> gpurun_out/act_bench_$w.err
  python - <<P
import json
d=json.loads(open("gpurun_out/act_bench_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["unit"], d["ms_per_step"], d.get("e2e",{}).get("value"))
P
done
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/act_launches.csv python bench.py --ncu > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/act_launches.csv 2>/dev/null | head -8
