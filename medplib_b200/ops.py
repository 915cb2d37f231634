"""Torch-tensor front end of the C ABI: validates shapes/dtypes, passes raw device pointers and the current stream.

Every function here enqueues hand-written sm_100a kernels from ``libmedplib_b200.so`` on torch's current CUDA
stream. There is no eager / CPU fallback: a CPU tensor or a missing library raises :class:`MplError`.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_GELU, ACT_QUICK_GELU, ACT_RELU, ACT_SILU, ACT_SIGMOID, DT_BF16, DT_F32,  # noqa: F401
                   MplError)

_ACT = {None: ACT_NONE, "none": ACT_NONE, "gelu": ACT_GELU, "quick_gelu": ACT_QUICK_GELU, "relu": ACT_RELU,
        "silu": ACT_SILU, "sigmoid": ACT_SIGMOID}
bf16 = torch.bfloat16


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _req(t, dtype, name):
    if not t.is_cuda:
        raise MplError(f"{name} must be a CUDA tensor (medplib_b200 has no CPU path)")
    if t.dtype != dtype:
        raise MplError(f"{name} must be {dtype}, got {t.dtype}")


def _rows(x):
    """View x [..., K] as 2-D [M, K] with unit inner stride (copy only if needed)."""
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < x2.shape[1]):
        x2 = x2.contiguous()
    return x2


def linear(x, weight, bias=None, act=None, residual=None, weight2=None, row_scale=None, m_dev=None,
           out_dtype=bf16, out=None, tile_n=0, force=None, ln_weight=None, ln_eps=1e-5, lora=None, ext=None, dual_out=None, silu_bwd=None):
    """y = epilogue(x @ weight.T); x [..., K] bf16, weight [N, K] bf16 (nn.Linear layout).

    weight may be a list/tuple of 1..3 same-shape matrices sharing x in one launch (returns a list of outputs).
    weight2 selects the fused SiLU(x W^T) * (x W2^T). force in {None, "tc", "skinny"}. See mpl_gemm_bf16.
    lora: up to two fused rank-r up-projections [(u [M, r] bf16 / f32, b [N, r] bf16 contiguous, scale, matrix index)]:
    out[i] = bf16(out[i] + bf16(scale * bf16(u b^T))) in the epilogue (tensor-core path, bf16 output).
    ext: (ext_a bf16 [M, 64], ext_b bf16 [len(weight) * N (2 * N with weight2), 64]): one extra k-block, the adapters
    inside the accumulator (mpl_gemm_args.ext_a / ext_b).
    """
    lib = _lib.load()
    ws = list(weight) if isinstance(weight, (list, tuple)) else [weight]
    nb = len(ws)
    _req(x, bf16, "x")
    for w in ws:
        _req(w, bf16, "weight")
        assert w.shape == ws[0].shape and w.stride() == ws[0].stride() and w.stride(1) == 1
    K = x.shape[-1]
    N = ws[0].shape[0]
    assert ws[0].shape[1] == K
    x2 = _rows(x)
    M = x2.shape[0]
    if silu_bwd is not None:
        # the dgrad of down_proj with the SiLU(gate) * up backward in its epilogue: (g, u) bf16 [M, N] are rewritten in
        # place with (dg, du); dh itself is not stored
        assert nb == 1 and out is None and weight2 is None and residual is None
        outs = [silu_bwd[0]]
    elif out is None:
        outs = [torch.empty((M, N), dtype=out_dtype, device=x.device) for _ in range(nb)]
    else:
        outs = list(out) if isinstance(out, (list, tuple)) else [out]
        outs = [o.reshape(-1, N) if o.dim() != 2 else o for o in outs]
    biases = list(bias) if isinstance(bias, (list, tuple)) else [bias] * nb if bias is None else [bias]
    a = _lib.GemmArgs()
    a.A, a.lda = x2.data_ptr(), x2.stride(0)
    for i in range(nb):
        a.B[i] = ws[i].data_ptr()
        assert outs[i].stride(1) == 1 and outs[i].stride(0) == outs[0].stride(0)
        a.C[i] = outs[i].data_ptr()
        if biases[i] is not None:
            _req(biases[i], bf16, "bias")
            a.bias[i] = biases[i].data_ptr()
    a.ldb, a.ldc = ws[0].stride(0), outs[0].stride(0)
    if weight2 is not None:
        assert nb == 1 and weight2.shape == ws[0].shape and weight2.stride() == ws[0].stride()
        a.B2 = weight2.data_ptr()
    if residual is not None:
        r2 = residual.reshape(-1, N) if residual.dim() != 2 else residual
        _req(r2, bf16, "residual")
        assert r2.stride(1) == 1
        a.residual, a.ldr = r2.data_ptr(), r2.stride(0)
    if row_scale is not None:
        _req(row_scale, torch.float32, "row_scale")
        a.row_scale = row_scale.data_ptr()
    if m_dev is not None:
        _req(m_dev, torch.int32, "m_dev")
        a.m_dev = m_dev.data_ptr()
    a.M, a.N, a.K, a.nb = M, N, K, nb
    a.act = _ACT[act]
    a.out_dtype = DT_F32 if outs[0].dtype == torch.float32 else DT_BF16
    a.tile_n = tile_n
    if ln_weight is not None:  # fused LlamaRMSNorm prologue (streaming path, M <= 16)
        _req(ln_weight, bf16, "ln_weight")
        a.ln_weight, a.ln_eps = ln_weight.data_ptr(), ln_eps
    unfused = []
    if lora and (os.environ.get("MPL_LORA_FUSE", "1") == "0" or any(t[0].shape[-1] != 8 for t in lora)):
        # (A/B switch, or a rank the epilogue does not take: the adapters as their own passes over the output)
        unfused, lora = lora, None
    if lora:
        assert len(lora) <= 2 and weight2 is None and outs[0].dtype == bf16
        keep = []
        for t, (u, b, sc, mat) in enumerate(lora):
            r = u.shape[-1]
            assert u.is_contiguous() and b.is_contiguous() and u.shape == (M, r) and b.shape == (N, r) and b.dtype == bf16
            assert r == 8 and u.dtype in (bf16, torch.float32)
            a.lora_r = r
            a.lora_u[t], a.lora_b[t] = u.data_ptr(), b.data_ptr()
            a.lora_scale[t], a.lora_u_f32[t], a.lora_mat[t] = float(sc), int(u.dtype == torch.float32), int(mat)
            keep.append((u, b))
    if silu_bwd is not None:
        sg, su = silu_bwd
        assert sg.dtype == bf16 and su.dtype == bf16 and sg.shape == (M, N) and su.shape == (M, N)
        assert sg.stride(1) == 1 and su.stride(1) == 1 and sg.stride(0) == su.stride(0)
        a.silu_bwd_g, a.silu_bwd_u, a.ldc = sg.data_ptr(), su.data_ptr(), sg.stride(0)
        a.C[0] = None
    if dual_out is not None:  # (weight2 given) also keep gate(x), up(x): bf16 [M, N] views with the row pitch of `out`
        dg, du = dual_out
        assert weight2 is not None and dg.dtype == bf16 and du.dtype == bf16 and dg.stride(1) == 1 and du.stride(1) == 1
        assert dg.stride(0) == outs[0].stride(0) and du.stride(0) == outs[0].stride(0)
        a.dual_g, a.dual_u = dg.data_ptr(), du.data_ptr()
    if ext is not None:
        ea, eb = ext
        assert ea.dtype == bf16 and eb.dtype == bf16 and ea.is_contiguous() and eb.is_contiguous()
        assert ea.shape[0] >= M and ea.shape[1] == 64 and eb.shape == ((2 if weight2 is not None else nb) * N, 64)
        a.ext_a, a.ext_b = ea.data_ptr(), eb.data_ptr()
    fn = {None: lib.mpl_linear_bf16, "tc": lib.mpl_gemm_bf16, "skinny": lib.mpl_skinny_gemm_bf16}[force]
    _lib.check(fn(ctypes.byref(a), _stream()), "mpl_linear_bf16")
    for u, b, sc, mat in unfused:
        from . import train_ops
        train_ops.lora_up_add(outs[mat], u, b, sc)
    lead = x.shape[:-1]
    res = [o.reshape(*lead, N) if out is None else o for o in outs]
    return res if isinstance(weight, (list, tuple)) else res[0]


def rmsnorm(x, weight, eps, out=None):
    """LlamaRMSNorm (HF 4.31 rounding). x [..., D] bf16."""
    lib = _lib.load()
    _req(x, bf16, "x"); _req(weight, bf16, "weight")
    x2 = _rows(x)
    y = torch.empty_like(x2) if out is None else out.reshape(-1, x.shape[-1])
    _lib.check(lib.mpl_rmsnorm(_ptr(x2), ctypes.c_longlong(x2.stride(0)), _ptr(weight), _ptr(y),
                               ctypes.c_longlong(y.stride(0)), x2.shape[0], x2.shape[1], ctypes.c_float(eps),
                               _stream()), "mpl_rmsnorm")
    return y.reshape(x.shape)


def layernorm(x, weight, bias, eps, act=None, out=None):
    """nn.LayerNorm over the last dim (optionally followed by GELU). x [..., D] bf16."""
    lib = _lib.load()
    _req(x, bf16, "x"); _req(weight, bf16, "weight"); _req(bias, bf16, "bias")
    x2 = _rows(x)
    y = torch.empty_like(x2) if out is None else out.reshape(-1, x.shape[-1])
    _lib.check(lib.mpl_layernorm(_ptr(x2), ctypes.c_longlong(x2.stride(0)), _ptr(weight), _ptr(bias), _ptr(y),
                                 ctypes.c_longlong(y.stride(0)), x2.shape[0], x2.shape[1], ctypes.c_float(eps),
                                 _ACT[act], _stream()), "mpl_layernorm")
    return y.reshape(x.shape)


def pool_layernorm(x, weight, bias, t_out, eps):
    """AdaptiveAvgPool1d over tokens fused with LayerNorm: x [n, t_in, D] -> [n, t_out, D]."""
    lib = _lib.load()
    _req(x, bf16, "x")
    x = x.contiguous()
    n, t_in, D = x.shape
    y = torch.empty((n, t_out, D), dtype=bf16, device=x.device)
    _lib.check(lib.mpl_pool_layernorm(_ptr(x), _ptr(weight), _ptr(bias), _ptr(y), n, t_in, t_out, D,
                                      ctypes.c_float(eps), _stream()), "mpl_pool_layernorm")
    return y


def attention(q, k, v, scale, causal=False, kv_mask=None, rel_h=None, rel_w=None, tk_dev=None, out=None,
              scratch=None):
    """softmax(scale * q k^T + bias + masks) v.  q [B, Tq, H, d], k/v [B, Tk, H, d] (any strides, d contiguous).

    Returns o [B, Tq, H, d] (contiguous unless `out` is given). rel_h/rel_w: f32 [B*H, Tq, kh] / [B*H, Tq, kw].
    """
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, bf16, n)
        assert t.dim() == 4 and t.stride(3) == 1
    B, Tq, H, d = q.shape
    Tk = k.shape[1]
    o = torch.empty((B, Tq, H, d), dtype=bf16, device=q.device) if out is None else out
    a = _lib.AttnArgs()
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    for name, t in (("q_stride", q), ("k_stride", k), ("v_stride", v), ("o_stride", o)):
        arr = getattr(a, name)
        arr[0], arr[1], arr[2] = t.stride(0), t.stride(1), t.stride(2)
    a.B, a.H, a.Tq, a.Tk, a.head_dim = B, H, Tq, Tk, d
    a.scale, a.causal = scale, int(causal)
    if kv_mask is not None:
        assert kv_mask.dtype in (torch.uint8, torch.bool) and kv_mask.is_contiguous() and kv_mask.shape == (B, Tk)
        a.kv_mask = kv_mask.data_ptr()
    if rel_h is not None:
        _req(rel_h, torch.float32, "rel_h"); _req(rel_w, torch.float32, "rel_w")
        assert rel_h.is_contiguous() and rel_w.is_contiguous()
        a.rel_h, a.rel_w = rel_h.data_ptr(), rel_w.data_ptr()
        a.rel_kh, a.rel_kw = rel_h.shape[-1], rel_w.shape[-1]
    if tk_dev is not None:
        _req(tk_dev, torch.int32, "tk_dev")
        a.tk_dev = tk_dev.data_ptr()
    if scratch is not None:  # zero-initialised uint8/float buffer: split-K decode
        a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel() * scratch.element_size()
    _lib.check(lib.mpl_attention(ctypes.byref(a), _stream()), "mpl_attention")
    return o


def _ll(v):
    return ctypes.c_longlong(int(v))


def moe_capacity(S, E, capacity_factor, min_capacity, k=1):
    """DeepSpeed _capacity: max(ceil(S / E * cf [* 2 for top-2]), min_capacity)."""
    import math
    return max(int(math.ceil((S / E) * capacity_factor * (2 if k == 2 else 1))), int(min_capacity), 1)


def moe_route(h, wg, k, capacity, noise=None):
    """Router + slot assignment (mpl_moe_route). h bf16 [S,D], wg f32 [E,D]. Returns a dict of device tensors."""
    lib = _lib.load()
    _req(h, bf16, "h"); _req(wg, torch.float32, "wg")
    h2 = _rows(h)
    S, D = h2.shape
    E = wg.shape[0]
    dev = h.device
    out = dict(logits=torch.empty((S, E), dtype=torch.float32, device=dev),
               gates=torch.empty((S, E), dtype=torch.float32, device=dev),
               expert=torch.empty((S, k), dtype=torch.int32, device=dev),
               gate=torch.empty((S, k), dtype=torch.float32, device=dev),
               slot=torch.empty((S, k), dtype=torch.int32, device=dev),
               kept=torch.empty((E,), dtype=torch.int32, device=dev),
               exp_counts=torch.empty((E,), dtype=torch.int32, device=dev),
               l_aux=torch.empty((1,), dtype=torch.float32, device=dev))
    a = _lib.MoeRouteArgs()
    a.h, a.ldh, a.wg = h2.data_ptr(), h2.stride(0), wg.contiguous().data_ptr()
    if noise is not None:
        _req(noise, torch.float32, "noise")
        a.noise = noise.contiguous().data_ptr()
    a.S, a.D, a.E, a.k, a.capacity = S, D, E, k, capacity
    for name in ("logits", "gates", "expert", "gate", "slot", "kept", "exp_counts", "l_aux"):
        setattr(a, name, out[name].data_ptr())
    _lib.check(lib.mpl_moe_route(ctypes.byref(a), _stream()), "mpl_moe_route")
    return out


def moe_dispatch(h, slot, rows, out=None):
    """xperm[slot[s,j]] = h[s]; returns xperm bf16 [rows, D] (rows = E * capacity; unused rows keep what `out` held,
    uninitialised when it is allocated here)."""
    lib = _lib.load()
    h2 = _rows(h)
    S, D = h2.shape
    k = slot.shape[1]
    xperm = torch.empty((rows, D), dtype=bf16, device=h.device) if out is None else out
    _lib.check(lib.mpl_moe_dispatch(_ptr(h2), _ll(h2.stride(0)), _ptr(slot), _ptr(xperm), S, k, D, _stream()),
               "mpl_moe_dispatch")
    return xperm


def moe_combine(y, slot, gate, residual=None):
    """out[s] = residual[s] + bf16(sum_j bf16(gate[s,j]) * y[slot[s,j]])."""
    lib = _lib.load()
    S, k = slot.shape
    D = y.shape[-1]
    out = torch.empty((S, D), dtype=bf16, device=y.device)
    r2 = _rows(residual) if residual is not None else None
    _lib.check(lib.mpl_moe_combine(_ptr(y), _ptr(slot), _ptr(gate), _ptr(r2), _ll(r2.stride(0) if r2 is not None else D),
                                   _ptr(out), _ll(D), S, k, D, _stream()), "mpl_moe_combine")
    return out


def rope_kv(q, k, v, cos, sin, pos0=0, k_cache=None, v_cache=None, pos_dev=None):
    """In-place RoPE on q,k [B,T,H,d] (views into a fused buffer allowed; row stride shared) + KV-cache append."""
    lib = _lib.load()
    B, T, H, d = q.shape
    ld = q.stride(1)
    assert k.stride(1) == ld and (v is None or v.stride(1) == ld) and q.stride(0) == T * ld and q.stride(2) == d
    Tmax = k_cache.shape[2] if k_cache is not None else 0
    _lib.check(lib.mpl_rope_kv(_ptr(q), _ptr(k), _ptr(v), _ll(ld), _ptr(cos), _ptr(sin), _ptr(k_cache), _ptr(v_cache),
                               B, T, H, d, Tmax, pos0, _ptr(pos_dev), _stream()), "mpl_rope_kv")


def gather_rows(idx, table=None, feats=None, D=None, out=None):
    """out[r] = table[idx[r]] (idx>=0) | 0 (idx==-1) | feats[-idx-2] (idx<=-2)."""
    lib = _lib.load()
    _req(idx, torch.int32, "idx")
    src = table if table is not None else feats
    D = D or src.shape[-1]
    rows = idx.numel()
    if out is None:
        out = torch.empty((rows, D), dtype=bf16, device=idx.device)
    _lib.check(lib.mpl_gather_rows(_ptr(table), _ll(table.stride(0) if table is not None else D), _ptr(feats),
                                   _ll(feats.stride(0) if feats is not None else D), _ptr(idx), _ptr(out),
                                   _ll(out.stride(0)), rows, D, _stream()), "mpl_gather_rows")
    return out


def argmax(logits, out=None):
    lib = _lib.load()
    _req(logits, torch.float32, "logits")
    rows, V = logits.shape
    if out is None:
        out = torch.empty((rows,), dtype=torch.int64, device=logits.device)
    _lib.check(lib.mpl_argmax_f32(_ptr(logits), _ll(logits.stride(0)), rows, V, _ptr(out), _stream()), "mpl_argmax")
    return out


def im2col_patch(img, P, k_pad=None):
    lib = _lib.load()
    _req(img, bf16, "img")
    img = img.contiguous()
    B, C, H, W = img.shape
    k_pad = k_pad or C * P * P
    out = torch.empty((B * (H // P) * (W // P), k_pad), dtype=bf16, device=img.device)
    _lib.check(lib.mpl_im2col_patch(_ptr(img), _ptr(out), B, C, H, W, P, k_pad, _stream()), "mpl_im2col_patch")
    return out


def im2col_nhwc(x, kh, kw, stride, pad, gate=None):
    lib = _lib.load()
    _req(x, bf16, "x")
    x = x.contiguous()
    B, H, W, C = x.shape
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    out = torch.empty((B * Ho * Wo, kh * kw * C), dtype=bf16, device=x.device)
    _lib.check(lib.mpl_im2col_nhwc(_ptr(x), _ptr(gate), _ptr(out), B, H, W, C, kh, kw, stride, pad, _stream()),
               "mpl_im2col_nhwc")
    return out


def clip_embed(patch, cls, pos, B):
    lib = _lib.load()
    D = patch.shape[-1]
    n = patch.shape[0] // B
    out = torch.empty((B, n + 1, D), dtype=bf16, device=patch.device)
    _lib.check(lib.mpl_clip_embed(_ptr(patch), _ptr(cls), _ptr(pos), _ptr(out), B, n, D, _stream()), "mpl_clip_embed")
    return out


def sam_relpos(q, rel_pos_h, rel_pos_w, hh, ww):
    """q bf16 [B, hh*ww, H, d] (strided view ok) -> rel_h f32 [B*H, hh*ww, hh], rel_w f32 [B*H, hh*ww, ww]."""
    lib = _lib.load()
    B, T, H, d = q.shape
    rel_h = torch.empty((B * H, T, hh), dtype=torch.float32, device=q.device)
    rel_w = torch.empty((B * H, T, ww), dtype=torch.float32, device=q.device)
    _lib.check(lib.mpl_sam_relpos(_ptr(q), _ll(q.stride(0)), _ll(q.stride(1)), _ll(q.stride(2)), _ptr(rel_pos_h),
                                  _ptr(rel_pos_w), _ptr(rel_h), _ptr(rel_w), B, H, hh, ww, d, _stream()),
               "mpl_sam_relpos")
    return rel_h, rel_w


def col_mean(x):
    lib = _lib.load()
    x = x.contiguous()
    B, T, C = x.shape
    out = torch.empty((B, C), dtype=bf16, device=x.device)
    _lib.check(lib.mpl_col_mean(_ptr(x), _ptr(out), B, T, C, _stream()), "mpl_col_mean")
    return out


def convt4s2_col2im(cols, B, Hi, Wi, C, skip=None):
    lib = _lib.load()
    _req(cols, torch.float32, "cols")
    out = torch.empty((B, 2 * Hi, 2 * Wi, C), dtype=bf16, device=cols.device)
    _lib.check(lib.mpl_convt4s2_col2im(_ptr(cols), _ptr(skip), _ptr(out), B, Hi, Wi, C, _stream()),
               "mpl_convt4s2_col2im")
    return out


def add(a, b):
    """bf16(a + b); b bf16 or f32, broadcast over leading dims when smaller."""
    lib = _lib.load()
    _req(a, bf16, "a")
    a = a.contiguous(); b = b.contiguous()
    out = torch.empty_like(a)
    _lib.check(lib.mpl_add(_ptr(a), _ptr(b), int(b.dtype == torch.float32), _ptr(out), _ll(a.numel()), _ll(b.numel()),
                           _stream()), "mpl_add")
    return out


def bilinear_resize(x, size, out_dtype=bf16):
    """x bf16 [N, Hin, Win] (inner stride 1) -> [N, Hout, Wout], align_corners=False."""
    lib = _lib.load()
    _req(x, bf16, "x")
    assert x.dim() == 3 and x.stride(2) == 1
    N, Hin, Win = x.shape
    out = torch.empty((N, size[0], size[1]), dtype=out_dtype, device=x.device)
    _lib.check(lib.mpl_bilinear_resize(_ptr(x), _ll(x.stride(0)), _ll(x.stride(1)), Hin, Win, _ptr(out),
                                       DT_F32 if out_dtype == torch.float32 else DT_BF16, size[0], size[1], N,
                                       _stream()), "mpl_bilinear_resize")
    return out


def region_sample_mean(fmap, pts, h, w):
    """fmap bf16 [h*w, C]; pts f32 [P,2] (x,y) in [0,1] -> bf16 [C]."""
    lib = _lib.load()
    _req(fmap, bf16, "fmap")
    fmap = fmap.contiguous()
    C = fmap.shape[-1]
    pts = pts.contiguous().float()
    out = torch.empty((C,), dtype=bf16, device=fmap.device)
    _lib.check(lib.mpl_region_sample_mean(_ptr(fmap), _ptr(pts), pts.shape[0], h, w, C, _ptr(out), _stream()),
               "mpl_region_sample_mean")
    return out


def grouped_linear(xperm, weights, m_dev, rows_per_group, weights2=None, out=None, row_map=None, row_gate=None,
                   residual=None, a_row_map=None, tile_n=0, m_total_hint=0):
    """Per-expert linears in one launch (mpl_grouped_gemm_bf16). xperm bf16 [G*rows_per_group, K]; weights: list of
    G [N,K]; m_dev int32 [G]. With row_map/row_gate (+residual): fused MoE combine into `out` [S, N] (token rows).
    With a_row_map (slot -> token): `xperm` is the un-dispatched [S, K] token matrix (fused MoE dispatch)."""
    lib = _lib.load()
    G = len(weights)
    K = xperm.shape[-1]
    N = weights[0].shape[0]
    a = _lib.GroupedGemmArgs()
    a.A, a.lda, a.a_group_stride = xperm.data_ptr(), xperm.stride(0), rows_per_group * xperm.stride(0)
    for g in range(G):
        a.B[g] = weights[g].data_ptr()
        if weights2 is not None:
            a.B2[g] = weights2[g].data_ptr()
    a.ldb = weights[0].stride(0)
    if out is None:
        out = torch.empty((G * rows_per_group, N), dtype=bf16, device=xperm.device)
    a.C, a.ldc, a.c_group_stride = out.data_ptr(), out.stride(0), rows_per_group * out.stride(0)
    a.m_dev = m_dev.data_ptr()
    if row_map is not None:
        a.row_map, a.row_gate, a.map_group_stride = row_map.data_ptr(), row_gate.data_ptr(), rows_per_group
    if a_row_map is not None:
        a.a_row_map, a.map_group_stride = a_row_map.data_ptr(), rows_per_group
    if residual is not None:
        a.residual, a.ldr = residual.data_ptr(), residual.stride(0)
    a.groups, a.M, a.N, a.K = G, rows_per_group, N, K
    a.out_dtype = DT_BF16
    a.tile_n, a.m_total_hint = tile_n, m_total_hint
    _lib.check(lib.mpl_grouped_gemm_bf16(ctypes.byref(a), _stream()), "mpl_grouped_gemm_bf16")
    return out


def moe_route_small(x, wg, k, capacity, ln_weight=None, ln_eps=1e-5, noise=None):
    """One-launch decode MoE front end (S <= 64): RMSNorm + route + slots + dispatch. Returns the mpl_moe_route dict
    plus h, xperm [E*capacity, D], tok_of_slot, gate_of_slot."""
    lib = _lib.load()
    _req(x, bf16, "x"); _req(wg, torch.float32, "wg")
    x2 = _rows(x)
    S, D = x2.shape
    E = wg.shape[0]
    dev = x.device
    out = dict(logits=torch.empty((S, E), dtype=torch.float32, device=dev),
               gates=torch.empty((S, E), dtype=torch.float32, device=dev),
               expert=torch.empty((S, k), dtype=torch.int32, device=dev),
               gate=torch.empty((S, k), dtype=torch.float32, device=dev),
               slot=torch.empty((S, k), dtype=torch.int32, device=dev),
               kept=torch.empty((E,), dtype=torch.int32, device=dev),
               exp_counts=torch.empty((E,), dtype=torch.int32, device=dev),
               l_aux=torch.empty((1,), dtype=torch.float32, device=dev),
               h=torch.empty((S, D), dtype=bf16, device=dev),
               xperm=torch.zeros((E * capacity, D), dtype=bf16, device=dev),
               tok_of_slot=torch.full((E * capacity,), -1, dtype=torch.int32, device=dev),
               gate_of_slot=torch.zeros((E * capacity,), dtype=torch.float32, device=dev))
    a = _lib.MoeRouteArgs()
    a.wg = wg.contiguous().data_ptr()
    if noise is not None:
        a.noise = noise.contiguous().data_ptr()
    a.S, a.D, a.E, a.k, a.capacity = S, D, E, k, capacity
    for name in ("logits", "gates", "expert", "gate", "slot", "kept", "exp_counts", "l_aux"):
        setattr(a, name, out[name].data_ptr())
    _lib.check(lib.mpl_moe_route_small(ctypes.byref(a), _ptr(x2), _ll(x2.stride(0)), _ptr(ln_weight),
                                       ctypes.c_float(ln_eps), _ptr(out["h"]), _ll(D), _ptr(out["xperm"]),
                                       _ptr(out["tok_of_slot"]), _ptr(out["gate_of_slot"]), _stream()),
               "mpl_moe_route_small")
    return out


# ------------------------------------------------------------------------------------------------ GeoRegionSampler (f-4)
def geo_point_table(fmaps, img_of_region, pts, h, w, ld):
    """point_sample of GeoSampler.py:263-276 into a stage-0 point table. fmaps bf16 [n_img, h*w, C]; img_of_region int32
    [R]; pts f32 [R, P, 2] = (row / H, col / W) -> bf16 [R, P, ld] = [features | row / H | col / W | 0...]."""
    lib = _lib.load()
    _req(fmaps, bf16, "fmaps"); _req(pts, torch.float32, "pts"); _req(img_of_region, torch.int32, "img_of_region")
    assert fmaps.is_contiguous() and pts.is_contiguous() and fmaps.shape[1] == h * w
    R, P, _ = pts.shape
    C = fmaps.shape[-1]
    table = torch.empty((R, P, ld), dtype=bf16, device=fmaps.device)
    _lib.check(lib.mpl_geo_point_table(_ptr(fmaps), _ptr(img_of_region), _ptr(pts), R, P, h, w, C, _ptr(table), ld,
                                       _stream()), "mpl_geo_point_table")
    return table


def geo_fps(table, d, S, start):
    """farthest_point_sample (:59-80) over the coordinates of a point table [R, N, ld] -> int32 [R, S]."""
    lib = _lib.load()
    _req(table, bf16, "table"); _req(start, torch.int32, "start")
    R, N, ld = table.shape
    assert table.is_contiguous() and start.shape == (R,)
    out = torch.empty((R, S), dtype=torch.int32, device=table.device)
    xy = ctypes.c_void_p(table.data_ptr() + 2 * d)
    _lib.check(lib.mpl_geo_fps(xy, ctypes.c_longlong(ld), R, N, S, _ptr(start), _ptr(out), _stream()), "mpl_geo_fps")
    return out


def geo_knn(table, d, fps_idx, k):
    """knn_point (:124-136): the k nearest of the table's N points to every anchor -> int32 [R, S, k]."""
    lib = _lib.load()
    _req(table, bf16, "table"); _req(fps_idx, torch.int32, "fps_idx")
    R, N, ld = table.shape
    S = fps_idx.shape[1]
    assert table.is_contiguous() and fps_idx.is_contiguous()
    out = torch.empty((R, S, k), dtype=torch.int32, device=table.device)
    xy = ctypes.c_void_p(table.data_ptr() + 2 * d)
    _lib.check(lib.mpl_geo_knn(xy, ctypes.c_longlong(ld), R, N, S, k, _ptr(fps_idx), _ptr(out), _stream()), "mpl_geo_knn")
    return out


def geo_group(table, fps_idx, knn_idx):
    """:302-308 -> (a1 [R*S*k, ld] = local - anchor, a2 [R*S*k, 2*ld] with the anchor rows in its second half)."""
    lib = _lib.load()
    R, N, ld = table.shape
    _, S, k = knn_idx.shape
    a1 = torch.empty((R * S * k, ld), dtype=bf16, device=table.device)
    a2 = torch.empty((R * S * k, 2 * ld), dtype=bf16, device=table.device)
    _lib.check(lib.mpl_geo_group(_ptr(table), ld, R, N, S, k, _ptr(fps_idx), _ptr(knn_idx), _ptr(a1), _ptr(a2),
                                 _stream()), "mpl_geo_group")
    return a1, a2


def geo_ln_pool(y, R, S, k, weight, bias, eps, mode, table=None, d=None, fps_idx=None, ldo=None):
    """LayerNorm + pool over the k neighbours (:152-156, 317). y bf16 [R*S*k, D] -> [R, S, ldo]; with ``table`` the
    result is the next stage's point table (anchor coordinates + padding appended)."""
    lib = _lib.load()
    _req(y, bf16, "y"); _req(weight, bf16, "weight"); _req(bias, bf16, "bias")
    D = y.shape[-1]
    assert y.is_contiguous() and y.shape[0] == R * S * k
    ldo = D if ldo is None else ldo
    out = torch.empty((R, S, ldo), dtype=bf16, device=y.device)
    if table is not None:
        xy, ld_src, N = ctypes.c_void_p(table.data_ptr() + 2 * d), table.shape[2], table.shape[1]
    else:
        xy, ld_src, N = None, 0, 0
    _lib.check(lib.mpl_geo_ln_pool(_ptr(y), R, S, k, D, _ptr(weight), _ptr(bias), ctypes.c_float(eps),
                                   {"mean": 0, "max": 1}[mode], xy, ctypes.c_longlong(ld_src), N, _ptr(fps_idx),
                                   _ptr(out), ctypes.c_longlong(ldo), _stream()), "mpl_geo_ln_pool")
    return out
