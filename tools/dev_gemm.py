"""Dev check of the tcgen05 GEMM on a GPU box: correctness vs torch fp32 and timing. Not part of tests/."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medplib_b200 import ops

torch.manual_seed(0)
dev = "cuda"

def check(M, N, K, **kw):
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    ref = x.float() @ w.float().t()
    bias = res = w2 = None
    if kw.get("bias"):
        bias = torch.randn(N, device=dev).bfloat16(); ref = ref + bias.float()
    act = kw.get("act")
    if kw.get("dual"):
        w2 = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        g = ref.bfloat16().float(); u = (x.float() @ w2.float().t()).bfloat16().float()
        ref = torch.nn.functional.silu(g).bfloat16().float() * u
    if act == "gelu": ref = torch.nn.functional.gelu(ref.bfloat16().float())
    if act == "relu": ref = torch.relu(ref)
    if kw.get("res"):
        res = torch.randn(M, N, device=dev).bfloat16(); ref = ref.bfloat16().float() + res.float()
    out_dtype = torch.float32 if kw.get("f32") else torch.bfloat16
    y = ops.linear(x, w, bias=bias, act=act, residual=res, weight2=w2, out_dtype=out_dtype, tile_n=kw.get("tile_n", 0))
    torch.cuda.synchronize()
    err = (y.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (1e-4 if kw.get("f32") else 1e-2) * max(scale, 1.0)
    ok = err <= tol
    print(f"M={M} N={N} K={K} {kw} max_err={err:.3e} scale={scale:.2f} {'OK' if ok else 'FAIL'}", flush=True)
    return ok

def bench(M, N, K, tile_n=0, iters=20):
    x = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.linear(x, w, out=out, tile_n=tile_n)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): ops.linear(x, w, out=out, tile_n=tile_n)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    for _ in range(3): torch.matmul(x, w.t())
    s.record()
    for _ in range(iters): torch.matmul(x, w.t())
    e.record(); torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / iters
    print(f"bench M={M} N={N} K={K} tile_n={tile_n}: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s | cuBLAS {ms2:.3f} ms {2*M*N*K/ms2/1e9:.1f} TFLOP/s", flush=True)

ok = True
ok &= check(128, 256, 64, tile_n=256)
ok &= check(128, 128, 64, tile_n=128)
ok &= check(128, 256, 256, tile_n=256)
ok &= check(256, 512, 1024)
ok &= check(615, 4096, 4096)
ok &= check(615, 4096, 4096, f32=True)
ok &= check(100, 200, 72, bias=True)
ok &= check(577, 1024, 592, bias=True, act="gelu")
ok &= check(615, 11008, 4096, dual=True)
ok &= check(300, 4096, 11008, res=True)
ok &= check(33, 32267, 4096, f32=True)
ok &= check(2048, 4096, 4096, tile_n=128)
ok &= check(2048, 4096, 4096, tile_n=256)
print("ALL OK" if ok else "SOME FAILED", flush=True)
if ok:
    for shp in [(615, 12288, 4096), (615, 4096, 4096), (615, 4096, 11008), (4096, 4096, 4096), (8192, 8192, 8192), (5120, 11008, 4096)]:
        bench(*shp, tile_n=256); bench(*shp, tile_n=128)
