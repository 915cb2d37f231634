#!/bin/bash
# First gpurun of a round: everything that was written after the previous round's GPU budget ran out, plus the
# measurements the next optimisation step needs.  ~3 GPU-minutes.  Usage:
#   gpurun --timeout 420 -- 'bash tools/first_gpu_call.sh'
# Outputs land in gpurun_out/first_call_*.{log,json}.
set -u
mkdir -p gpurun_out
# 1. tests that exist but have never run on a B200 (serving loop vs evaluate(), ICL encoder masks through the kernel)
MPL_RUN_UNVALIDATED=1 timeout 120 python -m pytest tests/test_zz_serve_gpu.py -m gpu -q -p no:cacheprovider --tb=short \
  > gpurun_out/first_call_unvalidated.log 2>&1
tail -5 gpurun_out/first_call_unvalidated.log
# 2. the division-free preprocess loop structure (candidate): bit-exactness, then the bench line next to the default's
MPL_PREPROCESS_V4=1 timeout 90 python -m pytest tests/test_preprocess_gpu.py -m gpu -q -p no:cacheprovider --tb=short \
  > gpurun_out/first_call_preprocess_v4_tests.log 2>&1
tail -3 gpurun_out/first_call_preprocess_v4_tests.log
timeout 60 python bench.py --workload preprocess --steps 50 --no-cpu-baseline > gpurun_out/first_call_preprocess_default.json 2>/dev/null
MPL_PREPROCESS_V4=1 timeout 60 python bench.py --workload preprocess --steps 50 --no-cpu-baseline \
  > gpurun_out/first_call_preprocess_v4.json 2>/dev/null
python - <<'PY'
import json
for n in ("default", "v4"):
    try:
        l = json.loads(open(f"gpurun_out/first_call_preprocess_{n}.json").read())
        print(n, "kernel us", round(l["roofline"]["kernel_ms_per_launch"] * 1e3, 1), "frac", round(l["roofline"]["frac"], 3))
    except Exception as e:
        print(n, "no line:", e)
PY
# 3. per-phase times of one layer of the persistent decode kernel at B = 1 (grounding) and B = 8 (VQA decode):
#    where the 112 us / layer at B = 1 go beyond the 61 us of weight streaming
for B in 1 8; do
  timeout 150 python tests/dev/dev_llama.py timing $B > gpurun_out/first_call_decode_phases_B$B.log 2>&1
  cat gpurun_out/first_call_decode_phases_B$B.log | tail -22
done
