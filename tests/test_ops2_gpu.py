"""GPU parity of the data-movement / MoE / SAM helper kernels against the CPU oracle. Index outputs are bit-exact."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


from parity import close as _close  # noqa: E402  (logs the measured error, asserts the stated tolerance)


@pytest.mark.parametrize("S,D,E,k,cf", [(615, 4096, 2, 1, 2.0), (8, 4096, 2, 1, 2.0), (5120, 256, 2, 1, 1.5),
                                        (300, 512, 4, 1, 1.0), (200, 256, 4, 2, 1.0), (64, 256, 2, 2, 2.0)])
def test_moe_route_dispatch_combine(dev, S, D, E, k, cf):
    from medplib_b200 import ops
    from oracle import moe
    g = torch.Generator().manual_seed(S + D + E)
    h = torch.randn(S, D, generator=g).to(bf16)
    wg = torch.randn(E, D, generator=g) * 0.5
    logits = h.float() @ wg.t()
    if k == 1:
        l_aux, gate, idx, slot, C, counts = moe.top1gating(logits, cf, 0)
        gate, idx, slot = gate[:, None], idx[:, None], slot[:, None]
    else:
        l_aux, gs, idxs, slots, C, counts = moe.top2gating(logits, cf, 0)
        gate, idx, slot = torch.stack(gs, 1), torch.stack(idxs, 1), torch.stack(slots, 1)
    assert C == ops.moe_capacity(S, E, cf, 0, k)
    r = ops.moe_route(h.to(dev), wg.to(dev), k, C)
    torch.cuda.synchronize()
    _close(r["logits"], logits, 1e-5, "logits")
    # ignore tokens whose top choices are numerically tied (fp32 summation order may flip them)
    srt = torch.sort(logits, dim=1, descending=True).values
    margin = srt[:, 0] - srt[:, 1] if (k == 1 or E == 2) else torch.minimum(srt[:, 0] - srt[:, 1], srt[:, 1] - srt[:, 2])
    ok = margin > 1e-3
    assert ok.float().mean() > 0.98
    if bool(ok.all()):
        assert torch.equal(r["expert"].cpu().long(), idx), "expert choice"
        ref_row = torch.where(slot >= 0, idx * C + slot, torch.full_like(slot, -1))
        assert torch.equal(r["slot"].cpu().long(), ref_row), "slots"
        assert torch.equal(r["exp_counts"].cpu().long(), counts), "exp_counts"
        _close(r["gate"], gate, 1e-5, "gate")
        _close(r["l_aux"], l_aux.reshape(1), 1e-5, "l_aux")
    # dispatch -> identity experts scaled per expert -> combine
    xperm = ops.moe_dispatch(h.to(dev), r["slot"], E * C)
    sl = r["slot"].cpu().long()
    for j in range(k):
        keep = sl[:, j] >= 0
        assert torch.equal(xperm.cpu()[sl[keep, j]], h[keep]), "dispatch rows"
    kept = r["kept"].cpu()
    for e in range(E):
        used = ((sl >= e * C) & (sl < (e + 1) * C)).sum().item()
        assert kept[e].item() == used
    y = torch.randn(E * C, D, generator=g).to(bf16)
    res = torch.randn(S, D, generator=g).to(bf16)
    out = ops.moe_combine(y.to(dev), r["slot"], r["gate"], res.to(dev))
    gt = r["gate"].cpu()
    acc = torch.zeros(S, D)
    for j in range(k):
        keep = sl[:, j] >= 0
        acc[keep] += gt[keep, j].to(bf16).float()[:, None] * y[sl[keep, j]].float()
    ref = acc.to(bf16).float() + res.float()
    _close(out, ref, 2 ** -7, "combine")


def test_moe_rts_overflow(dev):
    """Random-Token-Selection with an injected uniform sample: the kept set and slots match DeepSpeed's topk rule."""
    from medplib_b200 import ops
    from oracle import moe
    g = torch.Generator().manual_seed(3)
    S, D, E = 96, 256, 2
    h = torch.randn(S, D, generator=g).to(bf16)
    wg = torch.zeros(E, D)
    wg[0, :8] = 1.0  # unbalanced routing
    u = torch.rand(S, E, generator=g)
    logits = h.float() @ wg.t()
    l_aux, gate, idx, slot, C, counts = moe.top1gating(logits, 1.0, 0, rts_uniform=u)
    assert int(counts.max()) > C, "test must overflow an expert"
    r = ops.moe_route(h.to(dev), wg.to(dev), 1, C, noise=u.to(dev))
    ref_row = torch.where(slot >= 0, idx * C + slot, torch.full_like(slot, -1))
    assert torch.equal(r["slot"].cpu().long()[:, 0], ref_row)
    _close(r["gate"][:, 0], gate, 1e-5)


@pytest.mark.parametrize("B,T,H,d,pos0", [(1, 615, 32, 128, 0), (8, 1, 32, 128, 700), (2, 5, 2, 64, 3)])
def test_rope_kv(dev, B, T, H, d, pos0):
    from medplib_b200 import ops, engine
    from oracle import llama
    g = torch.Generator().manual_seed(T + d)
    qkv = torch.randn(B, T, 3, H, d, generator=g).to(bf16)
    Tmax = pos0 + T + 7
    cos, sin = llama.rope_tables(d, Tmax, 1e4, bf16)
    pos = torch.arange(pos0, pos0 + T)[None]
    q, k, v = qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2), qkv[:, :, 2].transpose(1, 2)
    qr, kr = llama.apply_rope(q, k, cos, sin, pos)
    dq = qkv.to(dev)
    kc = torch.zeros(B, H, Tmax, d, dtype=bf16, device=dev)
    vc = torch.zeros(B, H, Tmax, d, dtype=bf16, device=dev)
    c2, s2 = engine.rope_tables(d, Tmax, 1e4, dev)
    assert torch.equal(c2.cpu(), cos) and torch.equal(s2.cpu(), sin)
    pos_dev = torch.tensor([pos0], dtype=torch.int32, device=dev) if T == 1 else None
    ops.rope_kv(dq[:, :, 0], dq[:, :, 1], dq[:, :, 2], c2, s2, pos0=0 if pos_dev is not None else pos0, k_cache=kc,
                v_cache=vc, pos_dev=pos_dev)
    torch.cuda.synchronize()
    # same rounding points as the eager reference -> bit-exact
    assert torch.equal(dq[:, :, 0].cpu(), qr.transpose(1, 2)), "q"
    assert torch.equal(dq[:, :, 1].cpu(), kr.transpose(1, 2)), "k"
    assert torch.equal(kc[:, :, pos0:pos0 + T].cpu(), kr), "k cache"
    assert torch.equal(vc[:, :, pos0:pos0 + T].cpu(), v), "v cache"
    assert (kc[:, :, :pos0] == 0).all() and (kc[:, :, pos0 + T:] == 0).all()


def test_gather_rows_and_argmax(dev):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(0)
    table = torch.randn(100, 256, generator=g).to(bf16)
    feats = torch.randn(30, 256, generator=g).to(bf16)
    idx = torch.tensor([5, -1, -2, 99, -31, 0, -1, 7], dtype=torch.int32)
    out = ops.gather_rows(idx.to(dev), table.to(dev), feats.to(dev)).cpu()
    ref = torch.stack([table[5], torch.zeros(256).to(bf16), feats[0], table[99], feats[29], table[0],
                       torch.zeros(256).to(bf16), table[7]])
    assert torch.equal(out, ref)
    logits = torch.randn(8, 32267, generator=g)
    logits[3, 17] = logits[3, 30000] = 99.0  # tie -> lowest index
    assert torch.equal(ops.argmax(logits.to(dev)).cpu(), logits.argmax(-1))
    assert ops.argmax(logits.to(dev))[3].item() == 17


@pytest.mark.parametrize("C,S,P,kpad", [(3, 336, 14, 592), (3, 256, 16, 768), (3, 56, 14, 592)])
def test_im2col_patch(dev, C, S, P, kpad):
    from medplib_b200 import ops
    img = torch.randn(2, C, S, S, generator=torch.Generator().manual_seed(1)).to(bf16)
    out = ops.im2col_patch(img.to(dev), P, kpad).cpu()
    ref = F.unfold(img.float(), P, stride=P).transpose(1, 2).reshape(-1, C * P * P).to(bf16)
    assert torch.equal(out[:, :C * P * P], ref) and (out[:, C * P * P:] == 0).all()


@pytest.mark.parametrize("H,C,k,s,p,gate", [(16, 768, 3, 2, 1, True), (16, 256, 3, 1, 1, False), (21, 64, 3, 2, 1, False)])
def test_im2col_nhwc(dev, H, C, k, s, p, gate):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, H, H, C, generator=g).to(bf16)
    gt = torch.rand(2, C, generator=g).to(bf16) if gate else None
    out = ops.im2col_nhwc(x.to(dev), k, k, s, p, gate=gt.to(dev) if gate else None).cpu()
    xs = x.float() if not gate else (x.float() * gt.float()[:, None, None, :]).to(bf16).float()
    u = F.unfold(xs.permute(0, 3, 1, 2), k, stride=s, padding=p)  # [B, C*k*k, L] with (c, ky, kx) order
    L = u.shape[-1]
    ref = u.view(2, C, k * k, L).permute(0, 3, 2, 1).reshape(2 * L, k * k * C).to(bf16)
    assert torch.equal(out, ref)


def test_clip_embed_colmean_add(dev):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, n, D = 2, 576, 1024
    patch = torch.randn(B * n, D, generator=g).to(bf16)
    cls = torch.randn(D, generator=g).to(bf16)
    pos = torch.randn(n + 1, D, generator=g).to(bf16)
    out = ops.clip_embed(patch.to(dev), cls.to(dev), pos.to(dev), B).cpu()
    ref = torch.cat([cls.expand(B, 1, D), patch.view(B, n, D)], 1) + pos[None]
    assert torch.equal(out, ref)
    x = torch.randn(2, 256, 768, generator=g).to(bf16)
    _close(ops.col_mean(x.to(dev)), x.float().mean(1), 2 ** -8, "col_mean")
    a = torch.randn(256, 256, generator=g).to(bf16)
    pe = torch.randn(256, 256, generator=g)
    assert torch.equal(ops.add(a.to(dev), pe.to(dev)).cpu(), (a.float() + pe).to(bf16))
    row = torch.randn(256, generator=g).to(bf16)
    assert torch.equal(ops.add(a.to(dev), row.to(dev)).cpu(), a + row)


@pytest.mark.parametrize("hw,B,H", [(14, 4, 12), (16, 2, 12)])
def test_sam_relpos(dev, hw, B, H):
    from medplib_b200 import ops
    from oracle import sam
    g = torch.Generator().manual_seed(hw)
    d = 64
    qkv = torch.randn(B, hw * hw, 3, H, d, generator=g).to(bf16)
    rph = torch.randn(2 * hw - 1, d, generator=g).to(bf16)
    rpw = torch.randn(2 * hw - 1, d, generator=g).to(bf16)
    q = qkv[:, :, 0]
    rel_h, rel_w = ops.sam_relpos(qkv.to(dev)[:, :, 0], rph.to(dev), rpw.to(dev), hw, hw)
    Rh, Rw = sam._rel_pos(hw, hw, rph.float()), sam._rel_pos(hw, hw, rpw.float())
    rq = q.permute(0, 2, 1, 3).reshape(B * H, hw, hw, d).float()
    ref_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh).reshape(B * H, hw * hw, hw)
    ref_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw).reshape(B * H, hw * hw, hw)
    _close(rel_h, ref_h.to(bf16), 2 ** -7, "rel_h")
    _close(rel_w, ref_w.to(bf16), 2 ** -7, "rel_w")


def test_convt4s2_col2im(dev):
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, Hi, C = 2, 8, 64
    x = torch.randn(B, Hi, Hi, C, generator=g).to(bf16)
    w = (torch.randn(C, C, 4, 4, generator=g) * 0.1).to(bf16)  # ConvTranspose2d weight [Cin, Cout, 4, 4]
    skip = torch.randn(B, 2 * Hi, 2 * Hi, C, generator=g).to(bf16)
    w2 = w.permute(2, 3, 1, 0).reshape(16 * C, C)
    cols = ops.linear(x.to(dev).reshape(-1, C), w2.contiguous().to(dev), out_dtype=torch.float32, force="tc")
    out = ops.convt4s2_col2im(cols, B, Hi, Hi, C, skip=skip.to(dev)).cpu()
    y = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w.float(), stride=2, padding=1)
    ref = (F.relu(y.to(bf16).float()).permute(0, 2, 3, 1) + skip.float()).to(bf16)
    _close(out, ref, 2 ** -6, "convT col2im")


@pytest.mark.parametrize("size", [(336, 336), (300, 225), (64, 64), (17, 500)])
def test_bilinear_resize(dev, size):
    from medplib_b200 import ops
    x = torch.randn(2, 64, 64, generator=torch.Generator().manual_seed(5)).to(bf16)
    ref = F.interpolate(x.float()[None], size, mode="bilinear", align_corners=False)[0]
    out = ops.bilinear_resize(x.to(dev), size, out_dtype=torch.float32)
    _close(out, ref, 1e-5, "bilinear f32")
    out = ops.bilinear_resize(x.to(dev)[:, 3:51, 2:], size)  # cropped, strided view
    ref = F.interpolate(x[:, 3:51, 2:].float()[None], size, mode="bilinear", align_corners=False)[0]
    _close(out, ref.to(bf16), 2 ** -7, "bilinear bf16 cropped")


def test_region_sample_mean(dev):
    from medplib_b200 import ops
    from oracle import arch
    g = torch.Generator().manual_seed(6)
    C = 4096
    fmap = torch.randn(576, C, generator=g).to(bf16)
    masks = [(torch.rand(24, 24, generator=g) > 0.7).float(), torch.zeros(24, 24)]
    ref = arch.region_features(fmap[None], [masks], 512, bf16, bf16)[0]
    for m, r in zip(masks, ref):
        nz = (m.nonzero() / torch.tensor([[24, 24]])).to(bf16).float()  # the reference casts coords to the run dtype
        pts = nz.flip(1)
        out = ops.region_sample_mean(fmap.to(dev), pts.to(dev).reshape(-1, 2), 24, 24)
        _close(out, r, 2 ** -9, "region")  # half a bf16 ulp (measured: bit-exact once the sampling grid is rounded to bf16)


@pytest.mark.parametrize("M,N,K,nb", [(8, 4096, 4096, 3), (1, 4096, 4096, 1), (5, 512, 264, 1)])
def test_skinny_rmsnorm_prologue(dev, M, N, K, nb):
    """q,k,v straight from the un-normalised hidden state == rmsnorm kernel followed by the same GEMM (bit-exact)."""
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    x = (torch.randn(M, K, generator=g) * 2).to(bf16).to(dev)
    lnw = (1 + 0.1 * torch.randn(K, generator=g)).to(bf16).to(dev)
    ws = [(torch.randn(N, K, generator=g) * 0.05).to(bf16).to(dev) for _ in range(nb)]
    h = ops.rmsnorm(x, lnw, 1e-5)
    want = ops.linear(h, ws if nb > 1 else ws[0], force="skinny")
    got = ops.linear(x, ws if nb > 1 else ws[0], force="skinny", ln_weight=lnw, ln_eps=1e-5)
    for a, b in zip(got if nb > 1 else [got], want if nb > 1 else [want]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("S,D,F,E,k", [(8, 4096, 11008, 2, 1), (1, 4096, 11008, 2, 1), (6, 256, 512, 4, 2), (16, 256, 512, 2, 1)])
def test_moe_small_path_matches_general(dev, S, D, F, E, k):
    """mpl_moe_route_small + grouped expert GEMMs (+ fused combine) == rmsnorm + route + dispatch + per-expert GEMMs +
    combine, bit for bit."""
    from medplib_b200 import ops
    g = torch.Generator().manual_seed(S + D + E)
    x = (torch.randn(S, D, generator=g) * 2).to(bf16).to(dev)
    lnw = (1 + 0.1 * torch.randn(D, generator=g)).to(bf16).to(dev)
    wg = (torch.randn(E, D, generator=g) * 0.5).to(dev)
    wgate = [(torch.randn(F, D, generator=g) * D ** -0.5).to(bf16).to(dev) for _ in range(E)]
    wup = [(torch.randn(F, D, generator=g) * D ** -0.5).to(bf16).to(dev) for _ in range(E)]
    wdown = [(torch.randn(D, F, generator=g) * F ** -0.5).to(bf16).to(dev) for _ in range(E)]
    C = ops.moe_capacity(S, E, 2.0, 0, k)
    # general path
    h = ops.rmsnorm(x, lnw, 1e-5)
    r = ops.moe_route(h, wg, k, C)
    xp = ops.moe_dispatch(h, r["slot"], E * C)
    y = torch.zeros(E * C, D, dtype=bf16, device=dev)
    for e in range(E):
        h1 = ops.linear(xp[e * C:(e + 1) * C], wgate[e], weight2=wup[e], m_dev=r["kept"][e:e + 1])
        ops.linear(h1, wdown[e], m_dev=r["kept"][e:e + 1], out=y[e * C:(e + 1) * C])
    want = ops.moe_combine(y, r["slot"], r["gate"], x)
    # fused small path
    s = ops.moe_route_small(x, wg, k, C, ln_weight=lnw, ln_eps=1e-5)
    assert torch.equal(s["h"], h)
    for key in ("expert", "slot", "kept", "exp_counts"):
        assert torch.equal(s[key], r[key]), key
    assert torch.allclose(s["gate"], r["gate"], atol=1e-6) and torch.allclose(s["logits"], r["logits"], atol=1e-4)
    assert torch.allclose(s["l_aux"], r["l_aux"], atol=1e-6)
    h1 = ops.grouped_linear(s["xperm"], wgate, s["kept"], C, weights2=wup)
    if C <= 16:  # fused dispatch: the streaming kernel gathers its rows from h through the slot -> token map
        h1g = ops.grouped_linear(s["h"], wgate, s["kept"], C, weights2=wup, a_row_map=s["tok_of_slot"])
        for e in range(E):
            n = int(s["kept"][e])
            assert torch.equal(h1g[e * C:e * C + n], h1[e * C:e * C + n])
    if k == 1 and C <= 16:
        out = x.clone()
        ops.grouped_linear(h1, wdown, s["kept"], C, out=out, row_map=s["tok_of_slot"], row_gate=s["gate_of_slot"],
                           residual=out)
    else:
        y2 = ops.grouped_linear(h1, wdown, s["kept"], C)
        out = ops.moe_combine(y2, s["slot"], s["gate"], x)
    assert torch.equal(out, want)
