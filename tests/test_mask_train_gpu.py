"""GPU parity of the grounding-head backward kernels (medplib_b200/csrc/mask_train.cu, through the C ABI) against
torch.autograd in fp32 on the CPU. Tolerances: bf16 outputs within 2e-2 * max|ref|, fp32 accumulations within 5e-3."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _close(got, ref, rtol, name=""):
    got, ref = got.detach().float().cpu(), ref.detach().float()
    assert got.shape == ref.shape, f"{name}: {got.shape} vs {ref.shape}"
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale, f"{name}: max err {err:.4e} > {rtol} * scale {scale:.4e}"


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("M,N,K", [(7, 2048, 256), (256, 128, 256), (1024, 128, 64), (1, 4, 256), (2, 4096, 4096)])
def test_gemm_small_all_forms(dev, M, N, K):
    """The three uses: dX = dY W, dW += dY^T X (f32 accumulate), bf16 outer-product gradient."""
    from medplib_b200 import train_ops as T
    g = _g(M + N + K)
    dy = torch.randn(M, N, generator=g).to(bf16)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(bf16)
    x = torch.randn(M, K, generator=g).to(bf16)
    dx = T.gemm_small(dy.to(dev), W.to(dev))
    _close(dx, dy.float() @ W.float(), 2e-2, "dx")
    acc = torch.ones(N, K, device=dev)
    T.gemm_small(dy.to(dev), x.to(dev), out=acc, trans_a=True, accumulate=True)
    _close(acc, 1.0 + dy.float().t() @ x.float(), 5e-3, "dW accumulate")
    dw16 = T.gemm_small(dy.to(dev), x.to(dev), trans_a=True, out_dtype=bf16)
    _close(dw16, dy.float().t() @ x.float(), 2e-2, "dW bf16")
    yt = T.gemm_small(x.to(dev), W.to(dev), trans_b=True)
    _close(yt, x.float() @ W.float().t(), 2e-2, "x W^T")


def test_col_sum_and_accumulate(dev):
    from medplib_b200 import train_ops as T
    g = _g(1)
    X = torch.randn(37, 300, generator=g).to(bf16)
    out = torch.full((300,), 2.0, device=dev)
    T.col_sum(X.to(dev), out)
    _close(out, 2.0 + X.float().sum(0), 5e-3)
    v = torch.randn(300, generator=g)
    T.col_sum(v.to(dev), out)
    _close(out, 2.0 + X.float().sum(0) + v, 5e-3)


@pytest.mark.parametrize("rows,D,eps", [(7, 256, 1e-5), (1024, 64, 1e-6), (256, 256, 1e-5)])
def test_layernorm_bwd(dev, rows, D, eps):
    from medplib_b200 import train_ops as T
    g = _g(rows + D)
    x = torch.randn(rows, D, generator=g).to(bf16)
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(bf16)
    b = (0.1 * torch.randn(D, generator=g)).to(bf16)
    dy = torch.randn(rows, D, generator=g).to(bf16)
    xr, wr, br = x.float().requires_grad_(), w.float().requires_grad_(), b.float().requires_grad_()
    F.layer_norm(xr, (D,), wr, br, eps).backward(dy.float())
    dw = torch.zeros(D, device=dev)
    db = torch.zeros(D, device=dev)
    dx = T.layernorm_bwd(x.to(dev), w.to(dev), dy.to(dev), eps, dweight=dw, dbias=db)
    _close(dx, xr.grad, 2e-2, "dx")
    _close(dw, wr.grad, 5e-3, "dw")
    _close(db, br.grad, 5e-3, "db")


@pytest.mark.parametrize("act", ["gelu", "relu"])
def test_act_fwd_bwd(dev, act):
    from medplib_b200 import train_ops as T
    g = _g(3)
    x = (2 * torch.randn(1000, generator=g)).to(bf16)
    dy = torch.randn(1000, generator=g).to(bf16)
    xr = x.float().requires_grad_()
    y = F.gelu(xr) if act == "gelu" else F.relu(xr)
    y.backward(dy.float())
    _close(T.act_fwd(x.to(dev), act), y, 1e-2, "fwd")
    _close(T.act_bwd(x.to(dev), dy.to(dev), act), xr.grad, 1e-2, "bwd")


@pytest.mark.parametrize("Tq,Tk,H,d", [(6, 6, 8, 32), (6, 256, 8, 16), (256, 6, 8, 16), (7, 64, 2, 16)])
def test_attn_small_bwd(dev, Tq, Tk, H, d):
    """The mask decoder's three attention shapes (transformer.py:185-244): token self-attention, token->image,
    image->token."""
    from medplib_b200 import train_ops as T
    g = _g(Tq + Tk)
    C = H * d
    q = torch.randn(Tq, C, generator=g).to(bf16)
    k = torch.randn(Tk, C, generator=g).to(bf16)
    v = torch.randn(Tk, C, generator=g).to(bf16)
    do = torch.randn(Tq, C, generator=g).to(bf16)
    qr, kr, vr = (t.float().requires_grad_() for t in (q, k, v))
    s = (qr.view(Tq, H, d).transpose(0, 1) @ kr.view(Tk, H, d).transpose(0, 1).transpose(1, 2)) / math.sqrt(d)
    o = (torch.softmax(s, -1) @ vr.view(Tk, H, d).transpose(0, 1)).transpose(0, 1).reshape(Tq, C)
    o.backward(do.float())
    dq, dk, dv = T.attn_small_bwd(q.to(dev), k.to(dev), v.to(dev), do.to(dev), H, 1.0 / math.sqrt(d))
    _close(dq, qr.grad, 2e-2, "dq")
    _close(dk, kr.grad, 2e-2, "dk")
    _close(dv, vr.grad, 2e-2, "dv")


@pytest.mark.parametrize("hin,hout", [((64, 64), (336, 336)), ((64, 64), (70, 90)), ((64, 64), (40, 33))])
def test_bilinear_resize_bwd(dev, hin, hout):
    """Adjoint of F.interpolate(bilinear, align_corners=False) — up- and down-sampling."""
    from medplib_b200 import train_ops as T
    g = _g(hout[0])
    x = torch.randn(1, 1, *hin, generator=g, requires_grad=True)
    dy = torch.randn(1, *hout, generator=g).to(bf16)
    F.interpolate(x, hout, mode="bilinear", align_corners=False).backward(dy.float().unsqueeze(0))
    dx = T.bilinear_resize_bwd(dy.to(dev), hin)
    _close(dx, x.grad[0], 2e-2)


@pytest.mark.parametrize("shape", [(1, 70, 90), (1, 336, 336)])
def test_mask_losses_bwd(dev, shape):
    from medplib_b200 import train_ops as T
    from oracle import heads
    g = _g(shape[1])
    pred = (2 * torch.randn(shape, generator=g)).to(bf16)
    gt = (torch.rand(shape, generator=g) > 0.6).float()
    piou = torch.tensor([0.3]).to(bf16)
    w4 = torch.tensor([2.0, 0.5, 1.0, 1.5])
    pr, ir = pred.float().requires_grad_(), piou.float().requires_grad_()
    total = (w4[0] * heads.sigmoid_ce_loss(pr, gt, 1) + w4[1] * heads.dice_loss(pr, gt)
             + w4[2] * heads.mask_iou_loss(pr, gt, ir) + w4[3] * heads.focal_loss(pr, gt))
    total.backward()
    out, sums = T.mask_losses(pred.to(dev), gt.to(dev), piou.to(dev))
    dpred, dpi = T.mask_losses_bwd(pred.to(dev), gt.to(dev), piou.to(dev), sums, w4.to(dev))
    _close(dpred, pr.grad, 2e-2, "dpred")
    _close(dpi, ir.grad, 1e-2, "dpred_iou")


def test_attn_small_bwd_batched(dev):
    """Samples stacked along the row dimension (the B masks of a train step run as one batch)."""
    from medplib_b200 import train_ops as T
    B, Tq, Tk, H, d = 3, 6, 256, 8, 16
    g = _g(99)
    C = H * d
    q = torch.randn(B * Tq, C, generator=g).to(bf16)
    k = torch.randn(B * Tk, C, generator=g).to(bf16)
    v = torch.randn(B * Tk, C, generator=g).to(bf16)
    do = torch.randn(B * Tq, C, generator=g).to(bf16)
    qr, kr, vr = (t.float().requires_grad_() for t in (q, k, v))
    q4 = qr.view(B, Tq, H, d).permute(0, 2, 1, 3)
    k4 = kr.view(B, Tk, H, d).permute(0, 2, 1, 3)
    v4 = vr.view(B, Tk, H, d).permute(0, 2, 1, 3)
    o = (torch.softmax(q4 @ k4.transpose(-1, -2) / math.sqrt(d), -1) @ v4).permute(0, 2, 1, 3).reshape(B * Tq, C)
    o.backward(do.float())
    dq, dk, dv = T.attn_small_bwd(q.to(dev), k.to(dev), v.to(dev), do.to(dev), H, 1.0 / math.sqrt(d), batch=B)
    _close(dq, qr.grad, 2e-2, "dq")
    _close(dk, kr.grad, 2e-2, "dk")
    _close(dv, vr.grad, 2e-2, "dv")
