"""Dev tool (GPU box): time the tcgen05 GEMM at the four LLaMA-7B prefill shapes of the grounding step (T = 615) for
every compiled tile width, with the weights rotated over enough copies that none is L2-resident, and correctness of
every width against torch fp32.  python tools/gemm_tiles.py [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medplib_b200 import ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 615
dev = "cuda"
torch.manual_seed(0)
D, F = 4096, 11008
WIDTHS = [int(v) for v in os.environ["MPL_TILES"].split(",")] if os.environ.get("MPL_TILES") else [0] if (len(sys.argv) > 2 and sys.argv[2] == "ncu") else [0, 128, 144, 160, 176, 192, 208, 224, 240, 256, 1128, 1160, 1176, 1192, 1224, 1240, 1256]


def timeit(fn, n_rot, iters=24):
    for i in range(4):
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


x = rnd(T, D, scale=0.5)
R = 6
# o_proj
wo = [rnd(D, D) for _ in range(R)]
res = rnd(T, D, scale=0.5)
out = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
ref = (x.float() @ wo[0].float().t()).bfloat16().float() + res.float()
print(f"o_proj {T}x{D}x{D} (+residual), ideal {2*T*D*D/1375.4e6:.1f} us")
for w in WIDTHS:
    y = ops.linear(x, wo[0], residual=res, tile_n=w, force="tc")
    err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
    us = timeit(lambda i: ops.linear(x, wo[i], residual=res, out=out, tile_n=w, force="tc"), R)
    print(f"  tile_n {w:4d}: {us:7.1f} us  rel err {err:.2e}")
# qkv
wq = [[rnd(D, D) for _ in range(3)] for _ in range(R)]
outs = [torch.empty(T, D, device=dev, dtype=torch.bfloat16) for _ in range(3)]
refs = [(x.float() @ w_.float().t()) for w_ in wq[0]]
print(f"q,k,v {T}x3x{D}x{D}, ideal {2*T*3*D*D/1375.4e6:.1f} us")
for w in WIDTHS:
    ys = ops.linear(x, wq[0], tile_n=w, force="tc")
    err = max((y.float() - r).abs().max().item() / r.abs().max().item() for y, r in zip(ys, refs))
    us = timeit(lambda i: ops.linear(x, wq[i], out=outs, tile_n=w, force="tc"), R)
    print(f"  tile_n {w:4d}: {us:7.1f} us  rel err {err:.2e}")
# grouped gate/up (dual) and down, 2 experts
E = 2
cap = -(-T * 3 // (2 * E))  # ceil(T / E * 1.5)
kept = torch.tensor([T // 2, T - T // 2], device=dev, dtype=torch.int32)
xp = rnd(E * cap, D, scale=0.5)
wg_ = [[rnd(F, D) for _ in range(E)] for _ in range(3)]
wu_ = [[rnd(F, D) for _ in range(E)] for _ in range(3)]
h1 = torch.empty(E * cap, F, device=dev, dtype=torch.bfloat16)
print(f"gate|up grouped dual E={E} rows {kept.tolist()} x{F}x{D}, ideal {2*T*2*F*D/1375.4e6:.1f} us")
def ref_gu(e):
    n = int(kept[e])
    xe = xp[e * cap:e * cap + n].float()
    g = (xe @ wg_[0][e].float().t()).bfloat16().float()
    u = (xe @ wu_[0][e].float().t()).bfloat16().float()
    return torch.nn.functional.silu(g).bfloat16().float() * u
rg = [ref_gu(e) for e in range(E)]
for w in WIDTHS:
    y = ops.grouped_linear(xp, wg_[0], kept, cap, weights2=wu_[0], tile_n=w, m_total_hint=T)
    err = max((y[e * cap:e * cap + int(kept[e])].float() - rg[e]).abs().max().item() / rg[e].abs().max().item() for e in range(E))
    us = timeit(lambda i: ops.grouped_linear(xp, wg_[i], kept, cap, weights2=wu_[i], out=h1, tile_n=w, m_total_hint=T), 3)
    print(f"  tile_n {w:4d}: {us:7.1f} us  rel err {err:.2e}")
wd_ = [[rnd(D, F) for _ in range(E)] for _ in range(3)]
hh = rnd(E * cap, F, scale=0.5)
yo = torch.empty(E * cap, D, device=dev, dtype=torch.bfloat16)
rd = [(hh[e * cap:e * cap + int(kept[e])].float() @ wd_[0][e].float().t()) for e in range(E)]
print(f"down grouped E={E} x{D}x{F}, ideal {2*T*F*D/1375.4e6:.1f} us")
for w in WIDTHS:
    y = ops.grouped_linear(hh, wd_[0], kept, cap, tile_n=w, m_total_hint=T)
    err = max((y[e * cap:e * cap + int(kept[e])].float() - rd[e]).abs().max().item() / rd[e].abs().max().item() for e in range(E))
    us = timeit(lambda i: ops.grouped_linear(hh, wd_[i], kept, cap, out=yo, tile_n=w, m_total_hint=T), 3)
    print(f"  tile_n {w:4d}: {us:7.1f} us  rel err {err:.2e}")
if len(sys.argv) > 2 and sys.argv[2] == "ncu":
    # one launch of each shape at the automatic width, for `ncu -k regex:gemm_bf16 --set full` (the last 4 launches)
    torch.cuda.synchronize()
    ops.linear(x, wo[1], residual=res, out=out, force="tc")
    ops.linear(x, wq[1], out=outs, force="tc")
    ops.grouped_linear(xp, wg_[1], kept, cap, weights2=wu_[1], out=h1, m_total_hint=T)
    ops.grouped_linear(hh, wd_[1], kept, cap, out=yo, m_total_hint=T)
    torch.cuda.synchronize()
