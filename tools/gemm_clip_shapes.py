"""Dev tool (GPU box): the four GEMMs of a CLIP-L layer at M = 577 (one 336-px image) per tile width and epilogue
variant, weights rotated over enough copies that no panel is L2-resident.  python tools/gemm_clip_shapes.py [M]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medplib_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 577
dev = "cuda"
torch.manual_seed(0)
WIDTHS = [int(v) for v in os.environ.get("MPL_TILES", "0,64,96,128,144,160,192,224,256").split(",")]


def timeit(fn, n_rot, iters=48):
    for i in range(6):
        fn(i % n_rot)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for i in range(iters):
        fn(i % n_rot)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def rnd(*shape, scale=0.05):
    return (torch.randn(*shape, device=dev) * scale).bfloat16()


def run(name, N, K, nb=1, variants=("none", "bias", "bias+act", "bias+res")):
    R = max(4, int(200e6 // (N * K * 2 * nb)))
    x = rnd(M, K, scale=0.5)
    ws = [[rnd(N, K) for _ in range(nb)] for _ in range(R)]
    b = [rnd(N) for _ in range(nb)]
    res = rnd(M, N, scale=0.5)
    outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    print(f"{name}: {M}x{nb}x{N}x{K}, ideal {2 * M * nb * N * K / 1375.4e6:.1f} us, {R} weight copies")
    for var in variants:
        kw = {}
        if "bias" in var:
            kw["bias"] = b if nb > 1 else b[0]
        if "act" in var:
            kw["act"] = "quick_gelu"
        if "res" in var:
            kw["residual"] = res
        line = f"  {var:9s}"
        for w in WIDTHS:
            try:
                us = timeit(lambda i: ops.linear(x, ws[i] if nb > 1 else ws[i][0], out=outs if nb > 1 else outs[0],
                                                 tile_n=w, force="tc", **kw), R)
                line += f" | {w}: {us:6.1f}"
            except Exception as ex:  # a width the kernel does not take for this shape
                line += f" | {w}: n/a"
        print(line, flush=True)


run("qkv", 1024, 1024, nb=3, variants=("none", "bias"))
run("out_proj", 1024, 1024, variants=("none", "bias+res"))
run("fc1", 4096, 1024, variants=("none", "bias", "bias+act"))
run("fc2", 1024, 4096, variants=("none", "bias+res"))
