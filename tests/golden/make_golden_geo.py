"""Golden vectors for SURVEY §8 row f-4 (GeoRegionSampler), made by running the REFERENCE's own module
(model/rp_sampler/GeoSampler.py, imported from /root/reference in the authoring container).

Stored per case: the module's state dict, the inputs, the reference's output, and what its random / tie-dependent
steps did — the index tensors drawn by ``torch.randint`` / ``torch.randperm`` (recorded by wrapping those two
functions while the reference runs; the wrappers return what torch returns), the FPS indices and the kNN indices of
every stage (recorded by wrapping the module-level ``farthest_point_sample`` / ``knn_point``; behaviour unchanged).

Run:  python tests/golden/make_golden_geo.py        -> tests/golden/geo.pt
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
import importlib.util  # noqa: E402

# loaded by path: importing the ``model`` package would pull in deepspeed (absent here), which this file does not use
_spec = importlib.util.spec_from_file_location("ref_GeoSampler", "/root/reference/model/rp_sampler/GeoSampler.py")
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


def blob(rng, size, n_pix):
    """A connected region grown by a random walk (as the reference's unit test builds its masks, :363-388)."""
    m = np.zeros((size, size), dtype=np.int64)
    x, y = rng.integers(size), rng.integers(size)
    m[x, y] = 1
    for _ in range(n_pix):
        nb = [(x + dx, y + dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1)
              if (dx or dy) and 0 <= x + dx < size and 0 <= y + dy < size and m[x + dx, y + dy] == 0]
        if nb:
            x, y = nb[rng.integers(len(nb))]
            m[x, y] = 1
    return torch.from_numpy(m)


def run_case(seed, dtype, pooler, d, out_dim, n_init, subs, neighs, n_pix_list):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    mod = G.GeoRegionSampler(d, out_dim, n_init, subs, neighs, pooler_mode=pooler)
    for q in mod.parameters():  # LayerNorm affine away from (1, 0), everything bf16-representable
        q.data = (q.data + 0.05 * torch.randn_like(q)).to(torch.bfloat16).float()
    mod = mod.to(dtype).eval()
    n_img = len(n_pix_list)
    fmaps = (0.5 * torch.randn(n_img, 576, d)).to(torch.bfloat16).to(dtype)
    masks = [[blob(rng, 24, n) for n in per_img] for per_img in n_pix_list]
    draws, fps_rec, knn_rec = [], [], []
    o_randint, o_randperm, o_fps, o_knn = torch.randint, torch.randperm, G.farthest_point_sample, G.knn_point

    def randint(*a, **k):
        r = o_randint(*a, **k)
        draws.append(("randint", r.clone()))
        return r

    def randperm(*a, **k):
        r = o_randperm(*a, **k)
        draws.append(("randperm", r.clone()))
        return r

    def fps(xyz, npoint):
        r = o_fps(xyz, npoint)
        fps_rec.append(r.clone())
        return r

    def knn(nsample, xyz, new_xyz):
        r = o_knn(nsample, xyz, new_xyz)
        knn_rec.append(r.clone())
        return r

    torch.randint, torch.randperm, G.farthest_point_sample, G.knn_point = randint, randperm, fps, knn
    try:
        with torch.no_grad():
            out = mod([f for f in fmaps], masks, original_dtype=dtype, return_dtype=dtype)
    finally:
        torch.randint, torch.randperm, G.farthest_point_sample, G.knn_point = o_randint, o_randperm, o_fps, o_knn
    # the FPS starts are the LAST len(subs) randint draws of shape [R]; the ones before belong to rand_sample_repeat
    return dict(seed=seed, dtype=str(dtype), pooler=pooler, cfg=(d, out_dim, n_init, list(subs), list(neighs)),
                sd={k: v.to(torch.bfloat16) for k, v in mod.state_dict().items()},  # bf16-representable by construction
                fmaps=fmaps.to(torch.bfloat16), masks=[[m.to(torch.uint8) for m in per] for per in masks],
                draws=draws, fps=fps_rec, knn=knn_rec, out=out)


if __name__ == "__main__":
    cases = [
        # small regions (repeat-sampling), a large one (randperm), an image without regions
        run_case(1, torch.float32, "mean", 32, 48, 64, [16, 8], [6, 4], [[30, 90], [], [20]]),
        run_case(2, torch.float32, "max", 32, 48, 64, [16, 8], [6, 4], [[64], [10, 200]]),
        run_case(3, torch.bfloat16, "mean", 32, 48, 64, [16, 8], [6, 4], [[30, 90], [50]]),
        # the reference's own unit-test configuration (:351-356) at reduced width
        run_case(4, torch.float32, "max", 64, 80, 100, [50, 30], [20, 10], [[300], [520]]),
    ]
    path = os.path.join(HERE, "geo.pt")
    torch.save(cases, path)
    print(f"geo.pt: {os.path.getsize(path) / 1024:.1f} KiB")
