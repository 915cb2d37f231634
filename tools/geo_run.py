"""GeoRegionSampler at the configuration medplib_arch.py:136-141 names (CLIP-L features 1024 -> 4096, 512 initial points,
[128, 32] anchors, [24, 24] neighbours): one forward for an ncu capture (`--once`) or CUDA-event timing of N forwards.
python tools/geo_run.py [--regions 4] [--iters 20] [--once]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medplib_b200 import _lib  # noqa: E402
from medplib_b200.model.geo_sampler import GeoRegionSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--regions", type=int, default=4)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--once", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
bf16 = torch.bfloat16
torch.manual_seed(0)
mod = GeoRegionSampler(1024, 4096, 512, [128, 32], [24, 24], pooler_mode="max").to(device=dev, dtype=bf16).eval()
fm = (0.5 * torch.randn(1, 576, 1024)).to(bf16).to(dev)
masks = [[(torch.rand(24, 24) > 0.5).float() for _ in range(a.regions)]]
lib = _lib.load()
with torch.no_grad():
    for _ in range(1 if a.once else 3):
        out = mod(fm, masks, bf16, bf16)
    torch.cuda.synchronize()
    if not a.once:
        n0 = lib.mpl_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            out = mod(fm, masks, bf16, bf16)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(json.dumps({"workload": "GeoRegionSampler 1024->4096, 512 pts, [128,32]x[24,24]", "regions": a.regions,
                          "ms_per_forward": ms, "regions_per_s": a.regions / ms * 1e3,
                          "launches_per_forward": (lib.mpl_launch_count() - n0) / a.iters}))
assert out[0].shape == (a.regions, 4096) and bool(torch.isfinite(out[0].float()).all())
