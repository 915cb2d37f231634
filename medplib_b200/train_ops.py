"""Torch-tensor front end of the train-step entry points of the C ABI (include/medplib_b200.h, "Train step" section):
backward kernels of the LLaMA-MoE stack with LoRA, fused cross-entropy, mask losses, AdamW. Same rules as ops.py: CUDA
tensors only, kernels go on torch's current stream, no eager fallback."""
import ctypes
import os

import torch

from . import _lib
from .ops import _ll, _ptr, _req, _rows, _stream, bf16

f32 = torch.float32
_F = ctypes.c_float


def transpose(x, out=None, ld_out=None):
    """out[c, r] = x[r, c] for a 2-D bf16 tensor (inner stride 1). ld_out pads the output rows (extra columns zero)."""
    lib = _lib.load()
    _req(x, bf16, "x")
    assert x.dim() == 2 and x.stride(1) == 1
    R, C = x.shape
    if out is None:
        ld = ld_out or R
        out = torch.zeros((C, ld), dtype=bf16, device=x.device) if ld != R else torch.empty((C, R), dtype=bf16,
                                                                                           device=x.device)
    _lib.check(lib.mpl_transpose_bf16(_ptr(x), _ll(x.stride(0)), _ptr(out), _ll(out.stride(0)), R, C, _stream()),
               "mpl_transpose_bf16")
    return out


def lora_down(x, A, scale=1.0, out_f32=False, pad=None):
    """u = scale * x @ A.T; x bf16 [M,K], A bf16 [r,K] -> u [M,r] (bf16 or f32). pad = (u_pad bf16 [>= M, 64], col): a
    second bf16 copy of u in columns [col, col + r) of u_pad (the extension operand of the base GEMM)."""
    lib = _lib.load()
    _req(x, bf16, "x"); _req(A, bf16, "A")
    x2 = _rows(x)
    M, K = x2.shape
    r = A.shape[0]
    assert A.shape[1] == K and A.stride(1) == 1
    u = torch.empty((M, r), dtype=f32 if out_f32 else bf16, device=x.device)
    if pad is not None:
        buf, col = pad
        assert buf.dtype == bf16 and buf.is_contiguous() and buf.shape[0] >= M and buf.shape[1] == 64
        _lib.check(lib.mpl_lora_down_ext(_ptr(x2), _ll(x2.stride(0)), _ptr(A), _ll(A.stride(0)), _ptr(u), int(out_f32), M, K,
                                         r, _F(scale), _ptr(buf), int(col), _stream()), "mpl_lora_down_ext")
        return u
    _lib.check(lib.mpl_lora_down(_ptr(x2), _ll(x2.stride(0)), _ptr(A), _ll(A.stride(0)), _ptr(u), int(out_f32), M, K, r,
                                 _F(scale), _stream()), "mpl_lora_down")
    return u


def lora_up_add(y, u, Bm, scale=1.0, transposed=False):
    """y += scale * u @ Bm.T in place. y bf16 [M,N]; u [M,r] bf16/f32; Bm bf16 [N,r], or [r,N] when transposed."""
    lib = _lib.load()
    _req(y, bf16, "y"); _req(Bm, bf16, "Bm")
    assert y.dim() == 2 and y.stride(1) == 1 and u.is_contiguous()
    M, N = y.shape
    r = u.shape[1]
    if transposed:
        assert Bm.shape == (r, N)
        sn, sr = Bm.stride(1), Bm.stride(0)
    else:
        assert Bm.shape == (N, r)
        sn, sr = Bm.stride(0), Bm.stride(1)
    _lib.check(lib.mpl_lora_up_add(_ptr(y), _ll(y.stride(0)), _ptr(u), int(u.dtype == f32), _ptr(Bm), _ll(sn), _ll(sr),
                                   _F(scale), M, N, r, _stream()), "mpl_lora_up_add")
    return y


def rank_wgrad(X, U, out, scale=1.0, transposed=False):
    """out += scale * X.T @ U (f32 atomics). X bf16 [M,N]; U [M,r] bf16/f32; out f32 [N,r], or [r,N] when transposed."""
    lib = _lib.load()
    _req(X, bf16, "X"); _req(out, f32, "out")
    assert X.dim() == 2 and X.stride(1) == 1 and U.is_contiguous()
    M, N = X.shape
    r = U.shape[1]
    if transposed:
        assert out.shape == (r, N)
        sn, sr = out.stride(1), out.stride(0)
    else:
        assert out.shape == (N, r)
        sn, sr = out.stride(0), out.stride(1)
    _lib.check(lib.mpl_rank_wgrad(_ptr(X), _ll(X.stride(0)), _ptr(U), int(U.dtype == f32), _ptr(out), _ll(sn), _ll(sr),
                                  _F(scale), M, N, r, _stream()), "mpl_rank_wgrad")
    return out


def rmsnorm_bwd(x, weight, dy, eps, add=None, dweight=None, out=None):
    """dx = rmsnorm'(x; weight)·dy (+ add). x, dy, add bf16 [rows, D]; dweight f32 [D] accumulated when given."""
    lib = _lib.load()
    _req(x, bf16, "x"); _req(dy, bf16, "dy")
    x2, d2 = _rows(x), _rows(dy)
    a2 = _rows(add) if add is not None else None
    dx = torch.empty_like(x2) if out is None else out.reshape(-1, x.shape[-1])
    _lib.check(lib.mpl_rmsnorm_bwd(_ptr(x2), _ll(x2.stride(0)), _ptr(weight), _ptr(d2), _ll(d2.stride(0)), _ptr(a2),
                                   _ll(a2.stride(0) if a2 is not None else 0), _ptr(dx), _ll(dx.stride(0)),
                                   _ptr(dweight), x2.shape[0], x2.shape[1], _F(eps), _stream()), "mpl_rmsnorm_bwd")
    return dx.reshape(x.shape)


def silu_mul(g, u, out=None):
    lib = _lib.load()
    assert g.is_contiguous() and u.is_contiguous() and g.shape == u.shape
    h = torch.empty_like(g) if out is None else out
    _lib.check(lib.mpl_silu_mul(_ptr(g), _ptr(u), _ptr(h), _ll(g.numel()), _stream()), "mpl_silu_mul")
    return h


def silu_mul_bwd(g, u, dh):
    """In place: g <- dg, u <- du."""
    lib = _lib.load()
    assert g.is_contiguous() and u.is_contiguous() and dh.is_contiguous()
    _lib.check(lib.mpl_silu_mul_bwd(_ptr(g), _ptr(u), _ptr(dh), _ptr(g), _ptr(u), _ll(g.numel()), _stream()),
               "mpl_silu_mul_bwd")
    return g, u


def attention_fwd_lse(q, k, v, scale, causal=True, kv_mask=None):
    """ops.attention for training: also returns lse f32 [B*H, T] (log2 domain)."""
    lib = _lib.load()
    B, T, H, d = q.shape
    o = torch.empty((B, T, H, d), dtype=bf16, device=q.device)
    lse = torch.empty((B * H, T), dtype=f32, device=q.device)
    a = _lib.AttnArgs()
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    for name, t in (("q_stride", q), ("k_stride", k), ("v_stride", v), ("o_stride", o)):
        arr = getattr(a, name)
        arr[0], arr[1], arr[2] = t.stride(0), t.stride(1), t.stride(2)
    a.B, a.H, a.Tq, a.Tk, a.head_dim = B, H, T, k.shape[1], d
    a.scale, a.causal = scale, int(causal)
    if kv_mask is not None:
        assert kv_mask.dtype in (torch.uint8, torch.bool) and kv_mask.is_contiguous() and kv_mask.shape == (B, T)
        a.kv_mask = kv_mask.data_ptr()
    a.lse = lse.data_ptr()
    _lib.check(lib.mpl_attention(ctypes.byref(a), _stream()), "mpl_attention")
    return o, lse


def attention_bwd(q, k, v, o, d_o, lse, scale, dk, dv, causal=True, kv_mask=None):
    """Backward of the self-attention. q,k,v,o,d_o [B,T,H,d] (strided views fine; o and d_o share strides); dk, dv:
    bf16 [B,T,H,d] views to fill. Returns dq f32 [B,T,H,d]."""
    lib = _lib.load()
    B, T, H, d = q.shape
    assert o.stride() == d_o.stride()
    dq = torch.zeros((B, T, H, d), dtype=f32, device=q.device)
    delta = torch.empty((B * H, T), dtype=f32, device=q.device)
    a = _lib.AttnBwdArgs()
    a.q, a.k, a.v, a.o, a.d_o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr()
    for name, t in (("q_stride", q), ("k_stride", k), ("v_stride", v), ("o_stride", o), ("dk_stride", dk),
                    ("dv_stride", dv)):
        arr = getattr(a, name)
        arr[0], arr[1], arr[2] = t.stride(0), t.stride(1), t.stride(2)
    a.lse, a.delta, a.dq_f32, a.dk, a.dv = lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.B, a.H, a.T, a.head_dim, a.scale, a.causal = B, H, T, d, scale, int(causal)
    if kv_mask is not None:
        a.kv_mask = kv_mask.data_ptr()
    _lib.check(lib.mpl_attention_bwd(ctypes.byref(a), _stream()), "mpl_attention_bwd")
    return dq


def rope_bwd(dq_f32, dq, dk, cos, sin, pos0=0):
    """dq (bf16 view [B,T,H,d] of a fused buffer) = R(-pos) dq_f32; dk rotated in place."""
    lib = _lib.load()
    B, T, H, d = dq.shape
    ld = dq.stride(1)
    assert dk.stride(1) == ld and dq.stride(0) == T * ld and dq_f32.is_contiguous()
    _lib.check(lib.mpl_rope_bwd(_ptr(dq_f32), _ptr(dq), _ptr(dk), _ll(ld), _ptr(cos), _ptr(sin), B, T, H, d, pos0,
                                _stream()), "mpl_rope_bwd")


def moe_combine_bwd(dout, y, slot, gate, rows, C=None, kept=None):
    """Returns (dy bf16 [rows, D] zero elsewhere, dgate f32 [S,k]). With C / kept (rows per expert, kept count per expert)
    only the rows no kept slot writes are zeroed."""
    lib = _lib.load()
    d2 = _rows(dout)
    S, k = slot.shape
    D = y.shape[-1]
    dy = expert_buffer(rows, D, C, kept, y.device) if kept is not None else torch.zeros((rows, D), dtype=bf16,
                                                                                          device=y.device)
    dgate = torch.empty((S, k), dtype=f32, device=y.device)
    _lib.check(lib.mpl_moe_combine_bwd(_ptr(d2), _ll(d2.stride(0)), _ptr(y), _ptr(slot), _ptr(gate), _ptr(dy),
                                       _ptr(dgate), S, k, D, _stream()), "mpl_moe_combine_bwd")
    return dy, dgate


def moe_router_bwd(route, dgate, wg, dh, aux_scale=0.0):
    """top-1 / top-2 router backward; dh (bf16 [S,D]) is updated in place; returns dlogits f32 [S,E]."""
    lib = _lib.load()
    S, E = route["gates"].shape
    D = dh.shape[-1]
    dlogits = torch.empty((S, E), dtype=f32, device=dh.device)
    _lib.check(lib.mpl_moe_router_bwd(_ptr(route["gates"]), _ptr(route["expert"]), _ptr(route["slot"]), _ptr(dgate),
                                      _ptr(route["exp_counts"]), _F(aux_scale), _ptr(wg), _ptr(dlogits), _ptr(dh),
                                      _ll(dh.stride(0)), S, D, E, int(route["slot"].shape[1]), _stream()),
               "mpl_moe_router_bwd")
    return dlogits


def ce_fwd(logits, labels):
    """logits f32 [rows, V] (row stride free), labels i64 [rows] (<0 ignored) -> (lse [rows], acc [2] = loss sum, n)."""
    lib = _lib.load()
    _req(logits, f32, "logits"); _req(labels, torch.int64, "labels")
    rows, V = logits.shape
    lse = torch.empty((rows,), dtype=f32, device=logits.device)
    acc = torch.zeros((2,), dtype=f32, device=logits.device)
    _lib.check(lib.mpl_ce_fwd(_ptr(logits), _ll(logits.stride(0)), _ptr(labels), rows, V, _ptr(lse), _ptr(acc),
                              _stream()), "mpl_ce_fwd")
    return lse, acc


def ce_bwd(logits, labels, lse, acc, grad_out=None, ldd=None):
    """dlogits bf16 [rows, ldd] (ldd >= V, default V rounded up to 8; pad columns zero)."""
    lib = _lib.load()
    rows, V = logits.shape
    ldd = ldd or (V + 7) // 8 * 8
    dl = torch.empty((rows, ldd), dtype=bf16, device=logits.device)
    _lib.check(lib.mpl_ce_bwd(_ptr(logits), _ll(logits.stride(0)), _ptr(labels), rows, V, _ptr(lse), _ptr(acc),
                              _ptr(grad_out), _ptr(dl), _ll(ldd), _stream()), "mpl_ce_bwd")
    return dl


def scatter_add_rows(dx, idx, dtable=None, dfeats=None):
    lib = _lib.load()
    d2 = _rows(dx)
    _req(idx, torch.int32, "idx")
    _lib.check(lib.mpl_scatter_add_rows(_ptr(d2), _ll(d2.stride(0)), _ptr(idx), _ptr(dtable),
                                        _ll(dtable.stride(0) if dtable is not None else 0), _ptr(dfeats),
                                        _ll(dfeats.stride(0) if dfeats is not None else 0), d2.shape[0], d2.shape[1],
                                        _stream()), "mpl_scatter_add_rows")


def sumsq(g, out):
    lib = _lib.load()
    _req(g, f32, "g")
    _lib.check(lib.mpl_sumsq_f32(_ptr(g), _ll(g.numel()), _ptr(out), _stream()), "mpl_sumsq_f32")


def adamw(master, m, v, grad, param, lr, beta1, beta2, eps, weight_decay, step, sumsq_dev=None, max_norm=0.0,
          grad_scale=1.0):
    lib = _lib.load()
    assert master.is_contiguous() and grad.is_contiguous() and param.is_contiguous()
    _lib.check(lib.mpl_adamw(_ptr(master), _ptr(m), _ptr(v), _ptr(grad), _ptr(param), int(param.dtype == bf16),
                             _ll(master.numel()), _F(lr), _F(beta1), _F(beta2), _F(eps), _F(weight_decay), int(step),
                             _ptr(sumsq_dev), _F(max_norm), _F(grad_scale), _stream()), "mpl_adamw")


def adamw_multi(master, m, v, grad, chunks, lr, beta1, beta2, eps, weight_decay, step, sumsq_dev=None, max_norm=0.0,
                grad_scale=1.0):
    """One launch over the arena; chunks int64 [n, 4] on the device (see mpl_adamw_multi)."""
    lib = _lib.load()
    _req(chunks, torch.int64, "chunks")
    _lib.check(lib.mpl_adamw_multi(_ptr(master), _ptr(m), _ptr(v), _ptr(grad), _ptr(chunks), int(chunks.shape[0]), _F(lr),
                                   _F(beta1), _F(beta2), _F(eps), _F(weight_decay), int(step), _ptr(sumsq_dev),
                                   _F(max_norm), _F(grad_scale), _stream()), "mpl_adamw_multi")


def mask_losses(pred, gt, pred_iou):
    """pred bf16 [..] logits of ONE mask, gt f32 same numel, pred_iou bf16 scalar tensor -> (out4 f32, sums6 f32)."""
    lib = _lib.load()
    _req(pred, bf16, "pred"); _req(gt, f32, "gt")
    pred, gt = pred.contiguous(), gt.contiguous()
    out = torch.empty((4,), dtype=f32, device=pred.device)
    sums = torch.empty((6,), dtype=f32, device=pred.device)
    _lib.check(lib.mpl_mask_losses(_ptr(pred), _ptr(gt), _ptr(pred_iou), _ll(pred.numel()), _ptr(out), _ptr(sums),
                                   _stream()), "mpl_mask_losses")
    return out, sums


# ------------------------------------------------------------------------------------------ grounding-head backward
def _isf(t):
    return int(t.dtype == f32)


def gemm_small(A, Bm, out=None, trans_a=False, trans_b=False, out_dtype=bf16, accumulate=False):
    """C = op(A) @ op(B) with op = transpose when trans_*; A, B 2-D bf16 / f32 (any strides). `out` f32 + accumulate
    adds in place (weight gradients into the arena)."""
    lib = _lib.load()
    sam, sak = (A.stride(1), A.stride(0)) if trans_a else (A.stride(0), A.stride(1))
    M, K = (A.shape[1], A.shape[0]) if trans_a else A.shape
    sbk, sbn = (Bm.stride(1), Bm.stride(0)) if trans_b else (Bm.stride(0), Bm.stride(1))
    Kb, N = (Bm.shape[1], Bm.shape[0]) if trans_b else Bm.shape
    assert K == Kb, (A.shape, Bm.shape, trans_a, trans_b)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=A.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    _lib.check(lib.mpl_gemm_small(_ptr(A), _isf(A), _ll(sam), _ll(sak), _ptr(Bm), _isf(Bm), _ll(sbk), _ll(sbn), _ptr(out),
                                  _isf(out), _ll(out.stride(0)), int(accumulate), M, N, K, _stream()), "mpl_gemm_small")
    return out


def col_sum(X, out):
    """out (f32 [N]) += X.sum(0); X [M, N] bf16 / f32 (inner stride 1) or a vector."""
    lib = _lib.load()
    X2 = X.reshape(1, -1) if X.dim() == 1 else X
    assert X2.stride(1) == 1 and out.is_contiguous() and out.numel() == X2.shape[1]
    _lib.check(lib.mpl_col_sum(_ptr(X2), _isf(X2), _ll(X2.stride(0)), _ptr(out), X2.shape[0], X2.shape[1], _stream()),
               "mpl_col_sum")


def layernorm_bwd(x, weight, dy, eps, dweight=None, dbias=None):
    lib = _lib.load()
    x2, d2 = _rows(x), _rows(dy)
    dx = torch.empty_like(x2)
    _lib.check(lib.mpl_layernorm_bwd(_ptr(x2), _ll(x2.stride(0)), _ptr(weight), _ptr(d2), _ll(d2.stride(0)), _ptr(dx),
                                     _ll(dx.stride(0)), _ptr(dweight), _ptr(dbias), x2.shape[0], x2.shape[1], _F(eps),
                                     _stream()), "mpl_layernorm_bwd")
    return dx.reshape(x.shape)


def act_fwd(x, act):
    lib = _lib.load()
    x = x.contiguous()
    y = torch.empty_like(x)
    _lib.check(lib.mpl_act_fwd(_ptr(x), _ptr(y), _ll(x.numel()), _lib.ACT_GELU if act == "gelu" else _lib.ACT_RELU,
                               _stream()), "mpl_act_fwd")
    return y


def act_bwd(x, dy, act):
    lib = _lib.load()
    x, dy = x.contiguous(), dy.contiguous()
    dx = torch.empty_like(dy)
    _lib.check(lib.mpl_act_bwd(_ptr(x), _ptr(dy), _ptr(dx), _ll(x.numel()),
                               _lib.ACT_GELU if act == "gelu" else _lib.ACT_RELU, _stream()), "mpl_act_bwd")
    return dx


def attn_small_bwd(q, k, v, d_o, H, scale, batch=1):
    """q [B*Tq, H*d], k / v [B*Tk, H*d], d_o like q (inner stride 1, batches stacked) -> dq, dk, dv (bf16, contiguous)."""
    lib = _lib.load()
    C = q.shape[1]
    Tq, Tk = q.shape[0] // batch, k.shape[0] // batch
    d = C // H
    dq, dk, dv = torch.empty_like(q.contiguous()), torch.empty_like(k.contiguous()), torch.empty_like(v.contiguous())
    for t in (q, k, v, d_o):
        assert t.stride(1) == 1
    _lib.check(lib.mpl_attn_small_bwd(_ptr(q), _ll(q.stride(0)), _ptr(k), _ll(k.stride(0)), _ptr(v), _ll(v.stride(0)),
                                      _ptr(d_o), _ll(d_o.stride(0)), _ptr(dq), _ll(C), _ptr(dk), _ll(C), _ptr(dv), _ll(C),
                                      batch, Tq, Tk, H, d, _F(scale), _stream()), "mpl_attn_small_bwd")
    return dq, dk, dv


def bilinear_resize_bwd(dy, in_hw):
    """dy [N, Hout, Wout] (bf16 / f32, contiguous) -> dx bf16 [N, Hin, Win]."""
    lib = _lib.load()
    dy = dy.contiguous()
    N, Hout, Wout = dy.shape
    Hin, Win = in_hw
    dx = torch.empty((N, Hin, Win), dtype=bf16, device=dy.device)
    _lib.check(lib.mpl_bilinear_resize_bwd(_ptr(dy), _isf(dy), Hout, Wout, _ptr(dx), _ll(Hin * Win), _ll(Win), Hin, Win,
                                           N, _stream()), "mpl_bilinear_resize_bwd")
    return dx


def mask_losses_bwd(pred, gt, pred_iou, sums6, dloss4):
    """-> (dpred bf16 like pred, dpred_iou f32 [1])."""
    lib = _lib.load()
    pred, gt = pred.contiguous(), gt.contiguous()
    dpred = torch.empty_like(pred)
    dpi = torch.zeros((1,), dtype=f32, device=pred.device)
    _lib.check(lib.mpl_mask_losses_bwd(_ptr(pred), _ptr(gt), _ptr(pred_iou), _ptr(sums6), _ptr(dloss4.contiguous()),
                                       _ll(pred.numel()), _ptr(dpred), _ptr(dpi), _stream()), "mpl_mask_losses_bwd")
    return dpred, dpi


def mask_scale(x, mask, scale, out=None, accumulate=False):
    """out = (accumulate ? out : 0) + x * mask * scale; x / out bf16 contiguous, mask uint8 / bool same shape."""
    lib = _lib.load()
    _req(x, bf16, "x")
    assert x.is_contiguous() and mask.is_contiguous() and mask.numel() == x.numel()
    if out is None:
        out = torch.empty_like(x)
    assert out.is_contiguous()
    _lib.check(lib.mpl_mask_scale_bf16(_ptr(x), _ptr(mask), _F(scale), _ptr(out), int(accumulate), _ll(x.numel()),
                                       _stream()), "mpl_mask_scale_bf16")
    return out


def token_pool(x, t_out):
    """AdaptiveAvgPool1d over tokens: x bf16 [n, t_in, D] -> [n, t_out, D]."""
    lib = _lib.load()
    x = x.contiguous()
    n, t_in, D = x.shape
    y = torch.empty((n, t_out, D), dtype=bf16, device=x.device)
    _lib.check(lib.mpl_token_pool(_ptr(x), _ptr(y), n, t_in, t_out, D, _stream()), "mpl_token_pool")
    return y


def token_pool_bwd(dy, t_in):
    """Adjoint of AdaptiveAvgPool1d over tokens: dy bf16 [n, t_out, D] -> dx bf16 [n, t_in, D]."""
    lib = _lib.load()
    dy = dy.contiguous()
    n, t_out, D = dy.shape
    dx = torch.empty((n, t_in, D), dtype=bf16, device=dy.device)
    _lib.check(lib.mpl_token_pool_bwd(_ptr(dy), _ptr(dx), n, t_in, t_out, D, _stream()), "mpl_token_pool_bwd")
    return dx


def col2im_nhwc(dcols, B, H, W, C, kh, kw, stride, pad):
    """Adjoint of ops.im2col_nhwc: dcols bf16 [B*Ho*Wo, kh*kw*C] -> dx bf16 [B, H, W, C]."""
    lib = _lib.load()
    dcols = dcols.contiguous()
    dx = torch.empty((B, H, W, C), dtype=bf16, device=dcols.device)
    _lib.check(lib.mpl_col2im_nhwc(_ptr(dcols), _ptr(dx), B, H, W, C, kh, kw, stride, pad, _stream()), "mpl_col2im_nhwc")
    return dx


def expert_buffer(rows, width, C, kept, device):
    """bf16 [rows = E * C, width] whose rows past every expert's kept count are zero (the rest is left for the per-expert
    kernels to write)."""
    lib = _lib.load()
    if os.environ.get("MPL_ZERO_TAIL", "1") == "0":  # A/B switch: the whole buffer zero-filled, as before
        return torch.zeros((rows, width), dtype=bf16, device=device)
    buf = torch.empty((rows, width), dtype=bf16, device=device)
    _lib.check(lib.mpl_zero_tail_rows(_ptr(buf), _ll(width), rows // C, C, width, _ptr(kept), _stream()),
               "mpl_zero_tail_rows")
    return buf


def lora_pack(items_dev, n_items):
    """One launch fills the weight-side extension operands of every adapter (mpl_lora_pack; items: uint8 device tensor
    of 48-byte records)."""
    lib = _lib.load()
    _lib.check(lib.mpl_lora_pack(_ptr(items_dev), int(n_items), _stream()), "mpl_lora_pack")
