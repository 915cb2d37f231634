"""GPU parity of the individual C-ABI kernels against the CPU oracle (op level). Tolerances are stated per test:
bf16 outputs are compared at ~1 bf16 ulp of the output scale (2^-8 relative) unless noted."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16


from parity import close as _close  # noqa: E402  (logs the measured error, asserts the stated tolerance)


@pytest.mark.parametrize("M,N,K,kw", [
    (128, 256, 64, {}), (615, 4096, 4096, {}), (615, 4096, 4096, {"f32": True}), (100, 200, 72, {"bias": True}),
    (577, 1024, 592, {"bias": True, "act": "quick_gelu"}), (577, 4096, 1024, {"bias": True, "act": "gelu"}),
    (615, 11008, 4096, {"dual": True}), (300, 4096, 11008, {"res": True}), (33, 32267, 4096, {"f32": True}),
    (2048, 4096, 4096, {"tile_n": 128}), (2048, 4096, 4096, {"tile_n": 192}), (2048, 4096, 4096, {"tile_n": 256}),
    (615, 4096, 4096, {"nb": 3}), (577, 1024, 1024, {"nb": 3, "bias": True}), (256, 768, 192, {"act": "sigmoid"}),
    (700, 4096, 11008, {"res": True, "row_scale": True, "m_dev": 300}), (64, 768, 6912, {"act": "relu"}),
    # every compiled tile width (multiples of 16: the last 32-column chunk of a tile is partial for 144, 176, 208, 240;
    # dual halves 72 .. 120 columns), matrix edges that fall inside a partial chunk, fp32 output
    (615, 4096, 4096, {"tile_n": 144, "res": True}), (615, 4096, 4096, {"tile_n": 160, "bias": True}),
    (615, 4096, 4096, {"tile_n": 176, "res": True}), (615, 4096, 4096, {"tile_n": 208}),
    (615, 4096, 4096, {"tile_n": 224, "nb": 3}), (615, 4096, 4096, {"tile_n": 240, "f32": True}),
    (300, 1000, 512, {"tile_n": 144, "bias": True, "res": True}), (300, 1001, 520, {"tile_n": 176, "f32": True}),
    (615, 11008, 4096, {"dual": True, "tile_n": 240}), (615, 11008, 4096, {"dual": True, "tile_n": 144}),
    (615, 11008, 4096, {"dual": True, "tile_n": 208}), (330, 1000, 512, {"dual": True, "tile_n": 176}),
    # CTA-pair kernel (cta_group::2, 256-row tiles; tile_n = 1000 + width): ragged last pair (615 = 2 x 256 + 103, the
    # peer CTA of the last pair owns no valid row), M below one CTA, three matrices, every epilogue, dual, edges
    (615, 4096, 4096, {"tile_n": 1256}), (615, 4096, 4096, {"tile_n": 1176, "res": True}),
    (615, 4096, 4096, {"tile_n": 1224, "nb": 3}), (100, 200, 72, {"tile_n": 1128, "bias": True}),
    (2048, 4096, 4096, {"tile_n": 1256, "f32": True}), (577, 4096, 1024, {"tile_n": 1240, "bias": True, "act": "gelu"}),
    (615, 11008, 4096, {"dual": True, "tile_n": 1256}), (615, 11008, 4096, {"dual": True, "tile_n": 1240}),
    (700, 4096, 11008, {"tile_n": 1256, "res": True, "row_scale": True, "m_dev": 300}),
    (300, 1001, 520, {"tile_n": 1160, "f32": True}), (330, 1000, 512, {"dual": True, "tile_n": 1192}),
    (5112, 4096, 4096, {"tile_n": 1256}),
    # narrow tiles for small-N layers (CLIP fc2 / out_proj at M = 577: 40 tiles of 128 columns leave 108 SMs idle)
    (577, 1024, 4096, {"tile_n": 64, "bias": True, "res": True}), (577, 1024, 1024, {"tile_n": 96, "nb": 3, "bias": True}),
    (577, 1024, 4096, {"bias": True, "res": True}),
])
def test_gemm_tcgen05(dev, M, N, K, kw):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    nb = kw.get("nb", 1)
    x = (torch.randn(M, K, generator=g) * 0.5).to(bf16)
    ws = [(torch.randn(N, K, generator=g) * 0.05).to(bf16) for _ in range(nb)]
    bs = [torch.randn(N, generator=g).to(bf16) if kw.get("bias") else None for _ in range(nb)]
    w2 = (torch.randn(N, K, generator=g) * 0.05).to(bf16) if kw.get("dual") else None
    res = torch.randn(M, N, generator=g).to(bf16) if kw.get("res") else None
    rs = torch.rand(M, generator=g) if kw.get("row_scale") else None
    f32 = kw.get("f32", False)
    refs = []
    for w, b in zip(ws, bs):
        r = x.float() @ w.float().t()
        if b is not None:
            r = r + b.float()
        if w2 is not None:
            gte = r.to(bf16).float()
            up = (x.float() @ w2.float().t()).to(bf16).float()
            r = F.silu(gte).to(bf16).float() * up
        act = kw.get("act")
        if act:
            r = r if f32 else r.to(bf16).float()
            r = {"gelu": F.gelu, "relu": F.relu, "sigmoid": torch.sigmoid,
                 "quick_gelu": lambda t: t * torch.sigmoid(1.702 * t)}[act](r)
        if rs is not None:
            r = r.to(bf16).float() * rs[:, None]
        if res is not None:
            r = r.to(bf16).float() + res.float()
        refs.append(r)
    m_dev = None
    if "m_dev" in kw:
        m_dev = torch.tensor([kw["m_dev"]], dtype=torch.int32, device=dev)
    d = lambda t: t.to(dev) if t is not None else None
    out = None
    if m_dev is not None:
        out = torch.full((M, N), 7.0, dtype=bf16, device=dev)
    y = ops.linear(d(x), [d(w) for w in ws] if nb > 1 else d(ws[0]), bias=[d(b) for b in bs] if nb > 1 else d(bs[0]),
                   act=kw.get("act"), residual=d(res), weight2=d(w2), row_scale=d(rs), m_dev=m_dev, out=out,
                   out_dtype=torch.float32 if f32 else bf16, tile_n=kw.get("tile_n", 0), force="tc")
    ys = y if nb > 1 else [y]
    torch.cuda.synchronize()
    for yi, r in zip(ys, refs):
        if m_dev is not None:
            m = kw["m_dev"]
            assert (yi[m:].float() == 7.0).all(), "rows past *m_dev must not be written"
            yi, r = yi[:m], r[:m]
        _close(yi, r, 1e-4 if f32 else 1e-2, f"gemm {M}x{N}x{K} {kw}")


@pytest.mark.parametrize("M,N,K,kw", [
    (8, 4096, 4096, {}), (1, 4096, 4096, {"res": True}), (8, 11008, 4096, {"dual": True}),
    (8, 4096, 11008, {"res": True, "row_scale": True}), (8, 32267, 4096, {"f32": True}), (16, 256, 4096, {"bias": True}),
    (6, 2048, 256, {"bias": True, "act": "relu"}), (5, 200, 72, {"bias": True}), (4, 4096, 32, {}),
    (8, 11008, 4096, {"dual": True, "m_dev": 3}), (8, 4096, 4096, {"m_dev": 0}),
])
def test_skinny_gemm(dev, M, N, K, kw):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    x = (torch.randn(M, K, generator=g) * 0.5).to(bf16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(bf16)
    b = torch.randn(N, generator=g).to(bf16) if kw.get("bias") else None
    w2 = (torch.randn(N, K, generator=g) * 0.05).to(bf16) if kw.get("dual") else None
    res = torch.randn(M, N, generator=g).to(bf16) if kw.get("res") else None
    rs = torch.rand(M, generator=g) if kw.get("row_scale") else None
    f32 = kw.get("f32", False)
    r = x.float() @ w.float().t()
    if w2 is not None:
        r = F.silu(r.to(bf16).float()).to(bf16).float() * (x.float() @ w2.float().t()).to(bf16).float()
    if b is not None:
        r = r + b.float()
    if kw.get("act") == "relu":
        r = F.relu(r)
    if rs is not None:
        r = r.to(bf16).float() * rs[:, None]
    if res is not None:
        r = r.to(bf16).float() + res.float()
    d = lambda t: t.to(dev) if t is not None else None
    m_dev, out = None, None
    if "m_dev" in kw:
        m_dev = torch.tensor([kw["m_dev"]], dtype=torch.int32, device=dev)
        out = torch.full((M, N), 7.0, dtype=bf16, device=dev)
    y = ops.linear(d(x), d(w), bias=d(b), act=kw.get("act"), residual=d(res), weight2=d(w2), row_scale=d(rs),
                   m_dev=m_dev, out=out, out_dtype=torch.float32 if f32 else bf16, force="skinny")
    torch.cuda.synchronize()
    if m_dev is not None:
        m = kw["m_dev"]
        assert (y[m:].float() == 7.0).all()
        y, r = y[:m], r[:m]
        if m == 0:
            return
    _close(y, r, 1e-4 if f32 else 1e-2, f"skinny {M}x{N}x{K} {kw}")


@pytest.mark.parametrize("rows,D", [(615, 4096), (8, 4096), (3, 256)])
def test_rmsnorm(dev, rows, D):
    from medplib_b200 import ops
    from oracle import llama
    g = torch.Generator().manual_seed(rows + D)
    x = (torch.randn(rows, D, generator=g) * 2).to(bf16)
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(bf16)
    ref = llama.rmsnorm(x, w, 1e-5)
    y = ops.rmsnorm(x.to(dev), w.to(dev), 1e-5)
    # same rounding points as the reference: allow 1 bf16 ulp
    _close(y, ref, 2 ** -7, "rmsnorm")


@pytest.mark.parametrize("rows,D,act", [(577, 1024, None), (256, 768, None), (1024, 64, "gelu"), (6, 256, None),
                                        (40, 4096, None)])
def test_layernorm(dev, rows, D, act):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(rows + D)
    x = (torch.randn(rows, D, generator=g) * 2 + 0.3).to(bf16)
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(bf16)
    b = (0.1 * torch.randn(D, generator=g)).to(bf16)
    ref = F.layer_norm(x.float(), (D,), w.float(), b.float(), 1e-6)
    if act == "gelu":
        ref = F.gelu(ref.to(bf16).float())
    y = ops.layernorm(x.to(dev), w.to(dev), b.to(dev), 1e-6, act=act)
    _close(y, ref, 2 ** -7, "layernorm")


@pytest.mark.parametrize("n,t_in,t_out,D", [(2, 576, 256, 4096), (1, 441, 64, 256), (1, 12, 5, 64)])
def test_pool_layernorm(dev, n, t_in, t_out, D):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(t_in + D)
    x = torch.randn(n, t_in, D, generator=g).to(bf16)
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(bf16)
    b = (0.1 * torch.randn(D, generator=g)).to(bf16)
    pooled = F.adaptive_avg_pool1d(x.float().transpose(1, 2), t_out).transpose(1, 2).to(bf16)
    ref = F.layer_norm(pooled.float(), (D,), w.float(), b.float(), 1e-5)
    y = ops.pool_layernorm(x.to(dev), w.to(dev), b.to(dev), t_out, 1e-5)
    _close(y, ref, 2 ** -7, "pool_layernorm")


def _ref_attention(q, k, v, scale, causal=False, kv_mask=None, bias=None):
    # q [B,Tq,H,d] fp32 math; P rounded to bf16 before P.V like the reference's eager softmax(...).to(bf16) @ v
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * scale
    Tq, Tk = s.shape[-2:]
    if bias is not None:
        s = s + bias
    if causal:
        i = torch.arange(Tq)[:, None] + (Tk - Tq)
        s = s.masked_fill(torch.arange(Tk)[None, :] > i, float("-inf"))
    if kv_mask is not None:
        s = s.masked_fill(~kv_mask[:, None, None, :].bool(), float("-inf"))
    p = torch.softmax(s, -1)
    return (p @ vf).permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,H,Tq,Tk,d,causal,masked", [
    (1, 32, 615, 615, 128, True, False), (2, 4, 200, 200, 128, True, True), (1, 16, 577, 577, 64, False, False),
    (1, 8, 6, 256, 16, False, False), (1, 8, 256, 6, 16, False, False), (1, 8, 6, 6, 32, False, False),
    (2, 3, 70, 133, 64, False, False), (1, 2, 5, 77, 128, True, False),
    # tcgen05 path (attention_tc.cu): many key tiles, ragged tails, a causal window with a past (Tk > Tq), masks
    (2, 4, 639, 639, 128, True, True), (1, 4, 1299, 1299, 128, True, False), (2, 3, 130, 517, 128, True, True),
    (3, 16, 577, 577, 64, False, False), (2, 2, 300, 1000, 64, True, True), (1, 2, 33, 40, 128, False, False),
])
def test_attention_prefill(dev, B, H, Tq, Tk, d, causal, masked):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(Tq * 3 + Tk + d)
    q = torch.randn(B, Tq, H, d, generator=g).to(bf16)
    k = torch.randn(B, Tk, H, d, generator=g).to(bf16)
    v = torch.randn(B, Tk, H, d, generator=g).to(bf16)
    kv_mask = None
    if masked:
        kv_mask = torch.ones(B, Tk, dtype=torch.bool)
        kv_mask[1, Tk - 37:] = False
    scale = 1 / math.sqrt(d)
    ref = _ref_attention(q, k, v, scale, causal, kv_mask)
    o = ops.attention(q.to(dev), k.to(dev), v.to(dev), scale, causal=causal,
                      kv_mask=kv_mask.to(dev) if kv_mask is not None else None)
    _close(o, ref, 8e-3, "attention")  # 2 bf16 ulp (measured 0.97)


@pytest.mark.parametrize("B,H,hw,d", [(4, 12, 14, 64), (1, 12, 16, 64)])
def test_attention_relpos(dev, B, H, hw, d):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(hw)
    T = hw * hw
    q = torch.randn(B, T, H, d, generator=g).to(bf16)
    k = torch.randn(B, T, H, d, generator=g).to(bf16)
    v = torch.randn(B, T, H, d, generator=g).to(bf16)
    rel_h = torch.randn(B * H, T, hw, generator=g)
    rel_w = torch.randn(B * H, T, hw, generator=g)
    bias = (rel_h[:, :, :, None] + rel_w[:, :, None, :]).reshape(B, H, T, T)
    scale = d ** -0.5
    ref = _ref_attention(q, k, v, scale, bias=bias)
    o = ops.attention(q.to(dev), k.to(dev), v.to(dev), scale, rel_h=rel_h.to(dev), rel_w=rel_w.to(dev))
    _close(o, ref, 8e-3, "attention relpos")


@pytest.mark.parametrize("B,H,Tk,use_dev", [(8, 32, 615, False), (8, 32, 1127, True), (1, 32, 40, False), (2, 4, 3, True)])
def test_attention_decode(dev, B, H, Tk, use_dev):
    from medplib_b200 import ops
    d = 128
    g = torch.Generator().manual_seed(Tk)
    Tmax = Tk + 50
    q = torch.randn(B, 1, H, d, generator=g).to(bf16)
    kc = torch.randn(B, H, Tmax, d, generator=g).to(bf16)  # cache layout [B,H,Tmax,d]
    vc = torch.randn(B, H, Tmax, d, generator=g).to(bf16)
    scale = 1 / math.sqrt(d)
    ref = _ref_attention(q, kc[:, :, :Tk].permute(0, 2, 1, 3), vc[:, :, :Tk].permute(0, 2, 1, 3), scale)
    kd, vd = kc.to(dev), vc.to(dev)
    # split-K over the keys (scratch given) must agree with the single-CTA-per-head result
    scratch = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):  # twice: the counters are self-cleaning
        o2 = ops.attention(q.to(dev), kd[:, :, :Tk].permute(0, 2, 1, 3), vd[:, :, :Tk].permute(0, 2, 1, 3), scale,
                           scratch=scratch)
        _close(o2, ref, 8e-3, "decode attention split-K")
    if use_dev:
        tk_dev = torch.tensor([Tk], dtype=torch.int32, device=dev)
        o = ops.attention(q.to(dev), kd.permute(0, 2, 1, 3), vd.permute(0, 2, 1, 3), scale, tk_dev=tk_dev)
    else:
        o = ops.attention(q.to(dev), kd[:, :, :Tk].permute(0, 2, 1, 3), vd[:, :, :Tk].permute(0, 2, 1, 3), scale)
    _close(o, ref, 8e-3, "decode attention")
