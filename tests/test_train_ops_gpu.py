"""GPU parity of the train-step kernels (medplib_b200/csrc/train.cu) against plain fp32 PyTorch autograd on the CPU
(the oracle for floating-point backward kernels). Tolerances: bf16 outputs within a few bf16 ulps of the fp32 result
(rtol on max|ref| stated per test); fp32 accumulations within 2e-3."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _close(got, ref, rtol, name=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, f"{name}: shape {got.shape} vs {ref.shape}"
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale, f"{name}: max err {err:.4e} > {rtol} * scale {scale:.4e}"


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("R,C", [(64, 64), (100, 37), (4096, 11008), (33, 32267)])
def test_transpose(dev, R, C):
    from medplib_b200 import train_ops as T
    x = torch.randn(R, C, generator=_g(R + C)).to(bf16)
    out = T.transpose(x.to(dev))
    assert torch.equal(out.cpu(), x.t().contiguous())
    pad = T.transpose(x.to(dev), ld_out=(R + 7) // 8 * 8)
    assert torch.equal(pad.cpu()[:, :R], x.t().contiguous()) and bool((pad.cpu()[:, R:] == 0).all())


@pytest.mark.parametrize("M,K,N,r", [(615, 4096, 4096, 8), (77, 512, 1032, 16), (5112, 4096, 11008, 8)])
def test_lora_forward_and_backward(dev, M, K, N, r):
    """peft LoRA Linear: y = base + s * (x A^T) B^T and its gradients dA, dB, dx_lora."""
    from medplib_b200 import train_ops as T
    g = _g(M + K + N)
    x = (torch.randn(M, K, generator=g)).to(bf16)
    A = (torch.randn(r, K, generator=g) * 0.05).to(bf16)
    Bw = (torch.randn(N, r, generator=g) * 0.05).to(bf16)
    base = torch.randn(M, N, generator=g).to(bf16)
    dy = torch.randn(M, N, generator=g).to(bf16)
    s = 2.0
    xr, Ar, Br = x.float().requires_grad_(), A.float().requires_grad_(), Bw.float().requires_grad_()
    y_ref = base.float() + s * (xr @ Ar.t()) @ Br.t()
    y_ref.backward(dy.float())
    xd, Ad, Bd, dyd = x.to(dev), A.to(dev), Bw.to(dev), dy.to(dev)
    u = T.lora_down(xd, Ad)
    _close(u, x.float() @ A.float().t(), 1e-2, "u")
    y = T.lora_up_add(base.to(dev).clone(), u, Bd, s)
    _close(y, y_ref.detach(), 2e-2, "y")
    # backward
    Bt = T.transpose(Bd)
    du = T.lora_down(dyd, Bt, scale=s, out_f32=True)
    _close(du, s * dy.float() @ Bw.float(), 5e-3, "du")
    dB = torch.zeros(N, r, device=dev)
    T.rank_wgrad(dyd, u, dB, scale=s)
    _close(dB, Br.grad, 2e-2, "dB")
    dA = torch.zeros(r, K, device=dev)
    T.rank_wgrad(xd, du, dA, transposed=True)
    _close(dA, Ar.grad, 2e-2, "dA")
    dx = torch.zeros(M, K, dtype=bf16, device=dev)
    T.lora_up_add(dx, du, Ad, 1.0, transposed=True)
    _close(dx, xr.grad, 2e-2, "dx")


@pytest.mark.parametrize("rows,D", [(615, 4096), (33, 256)])
def test_rmsnorm_bwd(dev, rows, D):
    from medplib_b200 import train_ops as T
    from oracle import llama
    g = _g(rows + D)
    x = torch.randn(rows, D, generator=g).to(bf16)
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(bf16)
    dy = torch.randn(rows, D, generator=g).to(bf16)
    add = torch.randn(rows, D, generator=g).to(bf16)
    xr, wr = x.float().requires_grad_(), w.float().requires_grad_()
    llama.rmsnorm(xr, wr, 1e-5).backward(dy.float())
    dw = torch.zeros(D, device=dev)
    dx = T.rmsnorm_bwd(x.to(dev), w.to(dev), dy.to(dev), 1e-5, add=add.to(dev), dweight=dw)
    _close(dx, xr.grad + add.float(), 1.5e-2, "dx")
    _close(dw, wr.grad, 5e-3, "dw")
    dx2 = T.rmsnorm_bwd(x.to(dev), w.to(dev), dy.to(dev), 1e-5)
    _close(dx2, xr.grad, 1.5e-2, "dx (no add)")


def test_silu_mul_fwd_bwd(dev):
    from medplib_b200 import train_ops as T
    g = _g(5)
    a = (2 * torch.randn(300, 1024, generator=g)).to(bf16)
    b = torch.randn(300, 1024, generator=g).to(bf16)
    dh = torch.randn(300, 1024, generator=g).to(bf16)
    ar, br = a.float().requires_grad_(), b.float().requires_grad_()
    ref = F.silu(ar) * br
    ref.backward(dh.float())
    h = T.silu_mul(a.to(dev), b.to(dev))
    _close(h, ref.detach(), 1.5e-2, "h")
    ga, gb = a.to(dev).clone(), b.to(dev).clone()
    T.silu_mul_bwd(ga, gb, dh.to(dev))
    _close(ga, ar.grad, 1.5e-2, "dg")
    _close(gb, br.grad, 1.5e-2, "du")


def _attn_ref(q, k, v, scale, causal, kv_mask):
    """q,k,v fp32 [B,T,H,d] -> o [B,T,H,d] (eager softmax attention)."""
    B, T, H, d = q.shape
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    s = qh @ kh.transpose(2, 3) * scale
    mask = torch.zeros(B, 1, T, T)
    if causal:
        mask = mask + torch.full((T, T), float("-inf")).triu(1)
    if kv_mask is not None:
        mask = mask.masked_fill(~kv_mask[:, None, None, :].bool(), float("-inf"))
    p = torch.softmax(s + mask, dim=-1)
    return (p @ vh).permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,T,H,d,masked", [(1, 64, 2, 128, False), (2, 200, 3, 128, True), (1, 615, 2, 128, False),
                                            (2, 130, 2, 64, False)])
def test_attention_backward(dev, B, T, H, d, masked):
    """Self-attention (causal, optional right-padding key mask) backward against eager fp32 autograd. q/k/v live in a
    fused [B,T,3,H,d] buffer like the train engine's; dk/dv are written into the matching gradient buffer."""
    from medplib_b200 import train_ops as Tr
    g = _g(B * T + H + d)
    qkv = torch.randn(B, T, 3, H, d, generator=g).to(bf16)
    d_o = torch.randn(B, T, H, d, generator=g).to(bf16)
    kv_mask = None
    if masked:
        kv_mask = torch.ones(B, T, dtype=torch.bool)
        kv_mask[1, T - 37:] = False
    scale = 1.0 / math.sqrt(d)
    ref_in = qkv.float().requires_grad_()
    o_ref = _attn_ref(ref_in[:, :, 0], ref_in[:, :, 1], ref_in[:, :, 2], scale, True, kv_mask)
    valid = torch.ones(B, T, dtype=torch.bool) if kv_mask is None else kv_mask
    do_ref = d_o.float() * valid[:, :, None, None]  # padded query rows carry no gradient in the train step
    o_ref.backward(do_ref)
    qkv_d = qkv.to(dev)
    q, k, v = qkv_d[:, :, 0], qkv_d[:, :, 1], qkv_d[:, :, 2]
    km = kv_mask.to(torch.uint8).to(dev) if kv_mask is not None else None
    o, lse = Tr.attention_fwd_lse(q, k, v, scale, True, km)
    _close(o[valid.to(dev)], o_ref.detach()[valid], 2e-2, "o")
    dqkv = torch.zeros_like(qkv_d)
    dod = (d_o.float() * valid[:, :, None, None]).to(bf16).to(dev)
    dq = Tr.attention_bwd(q, k, v, o, dod, lse, scale, dqkv[:, :, 1], dqkv[:, :, 2], True, km)
    gref = ref_in.grad
    _close(dq, gref[:, :, 0], 2e-2, "dq")
    _close(dqkv[:, :, 1], gref[:, :, 1], 2e-2, "dk")
    _close(dqkv[:, :, 2], gref[:, :, 2], 2e-2, "dv")


def test_rope_backward(dev):
    from medplib_b200 import train_ops as Tr
    from medplib_b200.engine import rope_tables
    from oracle import llama
    B, T, H, d = 2, 50, 3, 128
    g = _g(9)
    dq = torch.randn(B, T, H, d, generator=g)
    dk = torch.randn(B, T, H, d, generator=g).to(bf16)
    cos, sin = llama.rope_tables(d, 64, 1e4, torch.float32)
    q0 = torch.zeros(B, H, T, d, requires_grad=True)
    k0 = torch.zeros(B, H, T, d, requires_grad=True)
    qr, kr = llama.apply_rope(q0, k0, cos.to(bf16).float(), sin.to(bf16).float(), torch.arange(T)[None])
    (qr * dq.permute(0, 2, 1, 3)).sum().backward(retain_graph=True)
    (kr * dk.float().permute(0, 2, 1, 3)).sum().backward()
    cd, sd = rope_tables(d, 64, 1e4, dev)
    buf = torch.zeros(B, T, 3, H, d, dtype=bf16, device=dev)
    buf[:, :, 1] = dk.to(dev)
    Tr.rope_bwd(dq.to(dev).contiguous(), buf[:, :, 0], buf[:, :, 1], cd, sd)
    _close(buf[:, :, 0], q0.grad.permute(0, 2, 1, 3), 1e-2, "dq")
    _close(buf[:, :, 1], k0.grad.permute(0, 2, 1, 3), 1e-2, "dk")


@pytest.mark.parametrize("S,D,E,cf,aux", [(300, 512, 2, 1.5, 0.0), (615, 4096, 2, 1.0, 0.01), (128, 256, 4, 2.0, 0.5)])
def test_moe_backward(dev, S, D, E, cf, aux):
    """Gradients of (sum(out * dout) + aux * l_aux) through the DeepSpeed top-1 MoE layer (oracle/moe.py autograd):
    d expert outputs, d router input, d wg."""
    from medplib_b200 import ops, train_ops as Tr
    from oracle import moe
    g = _g(S + D + E)
    h = torch.randn(S, D, generator=g).to(bf16)
    wg = torch.randn(E, D, generator=g) * 0.3
    dout = torch.randn(S, D, generator=g).to(bf16)
    C = moe.capacity(S, E, cf, 0)
    Ws = [(torch.randn(D, D, generator=g) / math.sqrt(D)) for _ in range(E)]
    hr, wgr = h.float().requires_grad_(), wg.clone().requires_grad_()
    seen = {}

    def expert(e):
        def f(t):
            y = t @ Ws[e].t()
            y.retain_grad()
            seen[e] = y
            return y
        return f
    out, l_aux, counts, logits = moe.moe_layer(hr[None], wgr, [expert(e) for e in range(E)], 1, cf, 0)
    (out[0] * dout.float()).sum().backward(retain_graph=True)
    (aux * l_aux).backward()
    # device
    hd, wgd = h.to(dev), wg.to(dev)
    route = ops.moe_route(hd, wgd, 1, C)
    xperm = ops.moe_dispatch(hd, route["slot"], E * C)
    y_dev = torch.zeros(E * C, D, dtype=bf16, device=dev)
    for e in range(E):
        y_dev[e * C:(e + 1) * C] = seen[e].detach().to(bf16).to(dev)
    dy, dgate = Tr.moe_combine_bwd(dout.to(dev), y_dev, route["slot"], route["gate"], E * C)
    kept = route["kept"].cpu()
    for e in range(E):
        n = int(kept[e])
        _close(dy[e * C:e * C + n], seen[e].grad[:n], 2e-2, f"dy expert {e}")
    # dispatch backward of an arbitrary upstream gradient on the expert inputs = combine with unit gates
    dxp = torch.randn(E * C, D, generator=g).to(bf16).to(dev)
    ones = torch.ones_like(route["gate"])
    dh_disp = ops.moe_combine(dxp, route["slot"], ones)
    slot, ex = route["slot"].cpu()[:, 0].long(), route["expert"].cpu()[:, 0].long()
    ref = torch.zeros(S, D)
    keep = slot >= 0
    ref[keep] = dxp.cpu().float()[slot[keep]]  # slot = global row e*C + position
    _close(dh_disp, ref, 1e-2, "dispatch backward")
    # router backward: dh (through the gate only: experts are constant w.r.t. h here) and dwg
    dh = torch.zeros(S, D, dtype=bf16, device=dev)
    dlogits = Tr.moe_router_bwd(route, dgate, wgd, dh, aux_scale=aux)
    dwg = torch.zeros(E, D, device=dev)
    Tr.rank_wgrad(hd, dlogits, dwg, transposed=True)
    _close(dwg, wgr.grad, 3e-2, "dwg")
    # reference dh through the gate path only = total dh minus the expert path (dispatch of W_e^T dy)
    dx_exp = torch.zeros(S, D)
    for e in range(E):
        sel = keep & (ex == e)
        dx_exp[sel] = (seen[e].grad @ Ws[e])[slot[sel] - e * C]
    _close(dh, hr.grad - dx_exp, 3e-2, "dh via router")


@pytest.mark.parametrize("rows,V", [(37, 300), (64, 32267)])
def test_cross_entropy(dev, rows, V):
    from medplib_b200 import train_ops as Tr
    g = _g(rows + V)
    logits = (3 * torch.randn(rows, V, generator=g))
    labels = torch.randint(0, V, (rows,), generator=g)
    labels[::5] = -100
    lr = logits.clone().requires_grad_()
    loss = F.cross_entropy(lr, labels, ignore_index=-100)
    (loss * 0.7).backward()
    ld, lab = logits.to(dev), labels.to(dev)
    lse, acc = Tr.ce_fwd(ld, lab)
    got = (acc[0] / acc[1]).item()
    assert abs(got - loss.item()) <= 1e-4 * abs(loss.item()), (got, loss.item())
    gout = torch.tensor([0.7], device=dev)
    dl = Tr.ce_bwd(ld, lab, lse, acc, gout)
    assert dl.shape[1] % 8 == 0 and bool((dl[:, V:] == 0).all())
    _close(dl[:, :V], lr.grad, 1e-2, "dlogits")


def test_scatter_add_rows(dev):
    from medplib_b200 import train_ops as Tr
    g = _g(3)
    V, nf, D, rows = 50, 7, 256, 200
    dx = torch.randn(rows, D, generator=g).to(bf16)
    idx = torch.randint(-nf - 1, V, (rows,), generator=g).to(torch.int32)
    dt, df = torch.zeros(V, D, device=dev), torch.zeros(nf, D, device=dev)
    Tr.scatter_add_rows(dx.to(dev), idx.to(dev), dt, df)
    rt, rf = torch.zeros(V, D), torch.zeros(nf, D)
    for r in range(rows):
        i = int(idx[r])
        if i >= 0:
            rt[i] += dx[r].float()
        elif i <= -2:
            rf[-i - 2] += dx[r].float()
    _close(dt, rt, 1e-5, "dtable")
    _close(df, rf, 1e-5, "dfeats")


def test_adamw_and_clip(dev):
    from medplib_b200 import train_ops as Tr
    g = _g(4)
    n = 10000
    p0 = torch.randn(n, generator=g)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.01)
    master, m, v = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    param = p0.to(bf16).to(dev)
    for step in range(1, 4):
        grad = torch.randn(n, generator=g) * 3
        ref_p.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        ss = torch.zeros(1, device=dev)
        Tr.sumsq(grad.to(dev), ss)
        assert abs(ss.item() - float((grad ** 2).sum())) <= 1e-3 * float((grad ** 2).sum())
        Tr.adamw(master, m, v, grad.to(dev), param, 3e-3, 0.9, 0.95, 1e-8, 0.01, step, ss, 1.0)
        _close(master, ref_p.detach(), 1e-5, f"step {step}")
        assert torch.equal(param.cpu(), master.cpu().to(bf16))


@pytest.mark.parametrize("shape", [(1, 70, 90), (1, 336, 336)])
def test_mask_losses_match_oracle(dev, shape):
    from medplib_b200 import train_ops as Tr
    from oracle import heads
    g = _g(sum(shape))
    pred = (2 * torch.randn(shape, generator=g)).to(bf16)
    gt = (torch.rand(shape, generator=g) > 0.7).float()
    iou = torch.tensor([0.4]).to(bf16)
    ref = [heads.sigmoid_ce_loss(pred.float(), gt, 1), heads.dice_loss(pred.float(), gt),
           heads.mask_iou_loss(pred.float(), gt, iou.float()), heads.focal_loss(pred.float(), gt)]
    out, sums = Tr.mask_losses(pred.to(dev), gt.to(dev), iou.to(dev))
    for i, name in enumerate(("bce", "dice", "iou", "focal")):
        assert abs(out[i].item() - float(ref[i])) <= 2e-4 * max(abs(float(ref[i])), 1e-3), (name, out[i].item(), ref[i])
