"""Torch-tensor front end of the C ABI: validates shapes/dtypes, passes raw device pointers and the current stream."""
import ctypes

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_GELU, ACT_QUICK_GELU, ACT_RELU, ACT_SILU, DT_BF16, DT_F32  # noqa: F401

_ACT = {None: ACT_NONE, "none": ACT_NONE, "gelu": ACT_GELU, "quick_gelu": ACT_QUICK_GELU, "relu": ACT_RELU,
        "silu": ACT_SILU}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _req(t, dtype, name):
    if not t.is_cuda:
        raise _lib.MplError(f"{name} must be a CUDA tensor (medplib_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.MplError(f"{name} must be {dtype}, got {t.dtype}")


def linear(x, weight, bias=None, act=None, residual=None, weight2=None, row_scale=None, m_dev=None,
           out_dtype=torch.bfloat16, out=None, tile_n=0):
    """y = epilogue(x @ weight.T); x [..., K] bf16, weight [N, K] bf16 (nn.Linear layout). See mpl_gemm_bf16."""
    lib = _lib.load()
    _req(x, torch.bfloat16, "x")
    _req(weight, torch.bfloat16, "weight")
    K = x.shape[-1]
    N = weight.shape[0]
    assert weight.shape[1] == K and weight.stride(1) == 1
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    a = _lib.GemmArgs()
    a.A, a.lda = x2.data_ptr(), x2.stride(0)
    a.B, a.ldb = weight.data_ptr(), weight.stride(0)
    a.B2 = weight2.data_ptr() if weight2 is not None else None
    if weight2 is not None:
        assert weight2.shape == weight.shape and weight2.stride() == weight.stride()
    a.C, a.ldc = out.data_ptr(), out.stride(0)
    a.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        r2 = residual.reshape(-1, N)
        _req(r2, torch.bfloat16, "residual")
        a.residual, a.ldr = r2.data_ptr(), r2.stride(0)
    a.row_scale = row_scale.data_ptr() if row_scale is not None else None
    a.m_dev = m_dev.data_ptr() if m_dev is not None else None
    a.M, a.N, a.K = M, N, K
    a.act = _ACT[act]
    a.out_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    a.tile_n = tile_n
    _lib.check(lib.mpl_gemm_bf16(ctypes.byref(a), _stream()), "mpl_gemm_bf16")
    return out.reshape(*x.shape[:-1], N)
