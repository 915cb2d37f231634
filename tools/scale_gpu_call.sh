#!/bin/bash
# The default bench line on N GPUs of one box, launched as the driver does:  gpurun --gpus N -- 'bash tools/scale_gpu_call.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --no-cpu-baseline > gpurun_out/final_n$N.json 2> gpurun_out/final_n$N.err
python - <<P
import json
d=json.loads(open("gpurun_out/final_n$N.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "grounding", round(d["value"],2), d["ms_per_step"])
for k,v in d["secondary"].items(): print(k, round(v["value"],2), v["unit"], round(v["ms_per_step"],2), v["config"].get("parallelism"))
P
