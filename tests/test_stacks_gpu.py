"""GPU parity of the native stack runners (LLaMA-MoE decoder, CLIP tower, SAM-Med2D encoder + mask decoder) against
the CPU oracle run in bf16 (the reference's eager cast points) on the same random weights and inputs.

Tolerances: both sides compute in bf16 with fp32 accumulation but in different summation orders and with fp32
softmax/LayerNorm internals on the GPU, so outputs agree to a few bf16 ulps of the activation scale; the bound used
is max|diff| <= 4e-2 * max|ref| (5 bf16 ulps) per stack, tighter where stated. Routing indices must match exactly on
tokens whose router margin exceeds the bf16 noise."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


from parity import close as _close  # noqa: E402  (logs the measured error, asserts the stated tolerance)


def _to(sd, dev):
    return {k: v.to(dev) for k, v in sd.items()}


LLAMA_CFGS = {
    "tiny_moe": dict(hidden_size=256, intermediate_size=512, num_layers=3, num_heads=2, vocab_size=64,
                     rms_norm_eps=1e-5, max_position_embeddings=256, rope_theta=1e4,
                     moe=dict(num_experts=2, top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0,
                              min_capacity=0, router_aux_loss_coef=0.01)),
    "tiny_dense": dict(hidden_size=256, intermediate_size=512, num_layers=2, num_heads=4, vocab_size=64,
                       rms_norm_eps=1e-5, max_position_embeddings=256, rope_theta=1e4, moe=None),
    "wide_moe_1layer": dict(hidden_size=4096, intermediate_size=11008, num_layers=1, num_heads=32, vocab_size=64,
                            rms_norm_eps=1e-5, max_position_embeddings=1024, rope_theta=1e4,
                            moe=dict(num_experts=2, top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0,
                                     min_capacity=0, router_aux_loss_coef=0.01)),
}


@pytest.mark.parametrize("name,B,T,steps,padded", [("tiny_moe", 2, 40, 3, True), ("tiny_dense", 1, 33, 2, False),
                                                   ("wide_moe_1layer", 1, 96, 2, False), ("tiny_moe", 8, 17, 4, False)])
def test_llama_stack_prefill_and_decode(dev, name, B, T, steps, padded):
    from medplib_b200 import engine
    from oracle import llama, weights
    cfg = LLAMA_CFGS[name]
    sd = weights.llama(cfg, seed=21)
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, T, cfg["hidden_size"], generator=g).to(bf16)
    am = torch.ones(B, T + steps, dtype=torch.bool)
    if padded:
        am[1, T - 9:T] = False  # right padding inside the prompt of sample 1
    ref = llama.model_forward(sd, cfg, x, am[:, :T])
    eng = engine.LlamaEngine(_to(sd, dev), cfg)
    cache = eng.new_cache(B, T + steps + 3)
    out = eng.forward(x.to(dev).clone(), cache, kv_mask=am.to(dev), want_hidden_states=True, want_router=True)
    torch.cuda.synchronize()
    valid = am[:, :T]
    margin_ok = valid.reshape(-1).clone()
    if cfg["moe"]:
        # a token whose router margin is within bf16 noise may legitimately flip experts; from that layer on it and,
        # through causal attention, every later token of the same sequence are excluded from the comparison. A flip
        # of a token with a clear margin is an error.
        for l, lg in enumerate(ref["gate_logits"]):
            got = out["gate_logits"][l].cpu()
            _close(got[margin_ok], lg[margin_ok], 1e-2 * (1 + 0.75 * l), f"router logits L{l}")  # bf16 noise grows with depth
            m = (lg[:, 0] - lg[:, 1]).abs()
            flip = (got.argmax(-1) != lg.argmax(-1)) & margin_ok
            assert bool((m[flip] <= 0.05 * lg.abs().max()).all()), f"layer {l}: expert flip despite a clear margin"
            margin_ok &= ~flip.reshape(B, T).cummax(dim=1).values.reshape(-1)
        assert margin_ok.float().mean() > 0.4
    keep = margin_ok.reshape(B, T)
    _close(out["last_hidden_state"].cpu()[keep], ref["last_hidden_state"][keep], 3.5e-2, "last hidden")
    for l in range(cfg["num_layers"]):
        _close(out["hidden_states"][l].cpu()[keep], ref["hidden_states"][l][keep], 2.5e-2, f"hidden {l}")
    if cfg["moe"] and bool(margin_ok.all()):
        for l in range(cfg["num_layers"]):
            assert torch.equal(out["exp_counts"][l].cpu().long(), ref["exp_counts"][l])
            _close(out["l_aux"][l], ref["moe_losses"][l], 1e-2, "l_aux")
    # KV cache contents
    kref = torch.stack([kv[0] for kv in ref["past_key_values"]])
    _close(cache.k[:, :, :, :T].cpu().permute(1, 3, 0, 2, 4)[keep], kref.permute(1, 3, 0, 2, 4)[keep], 2e-2, "k cache")
    # decode steps through the cache (fresh inputs per step; compares per-step hidden states)
    kv = ref["past_key_values"]
    seq_ok = keep.all(1) if padded is False else (keep | ~valid).all(1)
    for s in range(steps):
        xs = torch.randn(B, 1, cfg["hidden_size"], generator=g).to(bf16)
        r = llama.model_forward(sd, cfg, xs, am[:, :T + s + 1], kv)
        kv = r["past_key_values"]
        o = eng.forward(xs.to(dev).clone(), cache, kv_mask=am.to(dev), want_router=True)
        if cfg["moe"]:
            for l, lg in enumerate(r["gate_logits"]):
                got = o["gate_logits"][l].cpu()
                flip = (got.argmax(-1) != lg.argmax(-1)) & seq_ok
                m = (lg[:, 0] - lg[:, 1]).abs()
                assert bool((m[flip] <= 0.08 * lg.abs().max()).all()), f"decode step {s} layer {l}: flip, clear margin"
                seq_ok &= ~flip  # the sequence's cache now differs from the oracle's
        ok = seq_ok
        if ok.any():
            _close(o["last_hidden_state"].cpu()[ok], r["last_hidden_state"][ok], 2.5e-2, f"decode step {s}")
    assert cache.len == T + steps


def test_llama_training_capacity_drop(dev):
    """Training capacity factor with an overflowing expert and injected RTS uniforms: dropped tokens pass through the
    residual only (DeepSpeed semantics)."""
    from medplib_b200 import engine
    from oracle import llama, weights
    cfg = dict(LLAMA_CFGS["tiny_moe"], num_layers=1)
    cfg["moe"] = dict(cfg["moe"], capacity_factor=0.6)
    sd = weights.llama(cfg, seed=22)
    sd["model.layers.0.mlp.deepspeed_moe.gate.wg.weight"][1] = -sd["model.layers.0.mlp.deepspeed_moe.gate.wg.weight"][0]
    g = torch.Generator().manual_seed(1)
    B, T = 2, 32
    x = torch.randn(B, T, 256, generator=g).to(bf16)
    u = torch.rand(B * T, 2, generator=g)
    ref = llama.model_forward(sd, cfg, x, training=True, rts_uniforms=[u])
    eng = engine.LlamaEngine(_to(sd, dev), cfg)
    out = eng.forward(x.to(dev).clone(), eng.new_cache(B, T), training=True, moe_noise=[u.to(dev)], want_router=True)
    lg = ref["gate_logits"][0]
    ok = ((lg[:, 0] - lg[:, 1]).abs() > 0.05 * lg.abs().max()).reshape(B, T)
    _close(out["last_hidden_state"].cpu()[ok], ref["last_hidden_state"][ok], 1.5e-2, "train forward with drops")


CLIP_CFGS = {
    "tiny": dict(hidden_size=128, intermediate_size=256, num_layers=3, num_heads=2, image_size=56, patch_size=14),
    "vitl_3layers": dict(hidden_size=1024, intermediate_size=4096, num_layers=3, num_heads=16, image_size=336,
                         patch_size=14),
}


@pytest.mark.parametrize("name,B", [("tiny", 2), ("vitl_3layers", 1)])
def test_clip_stack(dev, name, B):
    from medplib_b200 import engine
    from oracle import clip, weights
    cfg = CLIP_CFGS[name]
    sd = weights.clip(cfg, seed=23)
    img = torch.randn(B, 3, cfg["image_size"], cfg["image_size"], generator=torch.Generator().manual_seed(2)).to(bf16)
    ref = clip.vision_tower(sd, "", img, cfg, select_layer=-2)
    eng = engine.ClipEngine(_to(sd, dev), cfg, "", select_layer=-2)
    out = eng.forward(img.to(dev))
    _close(out, ref, 1.7e-2, "clip features")


def test_sam_encoder_stack_vs_reference_golden(dev):
    """bf16 CUDA path vs the REFERENCE's own fp32 output (tests/golden/sam_encoder.pt)."""
    import inputs as gi
    from medplib_b200 import engine
    from oracle import weights
    g = torch.load(os.path.join(HERE, "golden", "sam_encoder.pt"), weights_only=False)
    cfg = gi.SAM_ENC_CFG
    sd = weights.sam_encoder(cfg, seed=gi.SAM_ENC_SEED, dtype=bf16)
    eng = engine.SamEncoderEngine(_to(sd, dev), cfg, "")
    out = eng.forward(gi.sam_encoder_images().to(dev))  # [B, 256, O] token-major
    ref = g["out"].permute(0, 2, 3, 1).reshape(out.shape)
    _close(out, ref, 2e-2, "sam encoder vs reference fp32")


@pytest.mark.parametrize("embed,heads,depth,B", [(128, 2, 4, 2), (768, 12, 3, 1)])
def test_sam_encoder_stack(dev, embed, heads, depth, B):
    from medplib_b200 import engine
    from oracle import sam, weights
    cfg = dict(embed_dim=embed, depth=depth, num_heads=heads, image_size=256, patch_size=16, out_chans=256)
    sd = weights.sam_encoder(cfg, seed=24)
    img = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(3)).to(bf16)
    ref = sam.image_encoder(sd, "", img, num_heads=heads)
    eng = engine.SamEncoderEngine(_to(sd, dev), cfg, "")
    out = eng.forward(img.to(dev))
    _close(out, ref.permute(0, 2, 3, 1).reshape(out.shape), 3e-2, "sam encoder")


def test_sam_mask_decoder_stack(dev):
    import inputs as gi
    from medplib_b200 import engine
    from oracle import sam, weights
    g = torch.load(os.path.join(HERE, "golden", "sam_head.pt"), weights_only=False)
    sd = weights.sam_head(seed=gi.SAM_HEAD_SEED)
    emb, text = gi.sam_head_inputs()
    emb, text = emb.to(bf16), text.to(bf16)
    dpe = sam.dense_pe(sd, "prompt_encoder.", (16, 16))
    sparse, dense = sam.prompt_encoder_text(sd, "prompt_encoder.", text, (16, 16))
    ref_m, ref_iou = sam.mask_decoder(sd, "mask_decoder.", emb, dpe, sparse.to(bf16), dense, multimask_output=False)
    eng = engine.MaskDecoderEngine(_to(sd, dev), "")
    _close(eng._keep[0], dpe[0].permute(1, 2, 0).reshape(256, 256), 1e-5, "dense pe")
    tok_major = emb[0].permute(1, 2, 0).reshape(256, 256).contiguous()
    mask, iou = eng.forward(tok_major.to(dev), text.reshape(-1).to(dev))
    _close(mask, ref_m, 2.5e-2, "low-res mask vs bf16 oracle")
    _close(iou, ref_iou, 4e-2, "iou vs bf16 oracle")
    _close(mask, g["masks"], 3e-2, "low-res mask vs reference fp32 golden")
    # mask indices (sigmoid > 0.1 <=> logit > log(1/9)) agree with the reference away from the threshold
    thr = -2.1972246
    far = (g["masks"] - thr).abs() > 0.05 * g["masks"].abs().max()
    assert torch.equal((mask.float().cpu() > thr)[far], (g["masks"] > thr)[far])


@pytest.mark.parametrize("name,B,T,steps,padded", [("tiny_moe", 8, 17, 5, False), ("tiny_moe", 3, 70, 4, True),
                                                   ("wide_moe_1layer", 1, 96, 3, False),
                                                   ("wide_moe_1layer", 8, 40, 2, False)])
def test_decode_kernel_matches_general_path(dev, name, B, T, steps, padded):
    """The one-kernel decode step (llama_decode.cu) against the per-op runner on the same engine and cache contents:
    the appended K/V rows are bit-identical (same RoPE roundings), hidden states agree to bf16 noise (the attention
    split-K partition differs, everything else has the same rounding points), routing decisions are identical."""
    from medplib_b200 import engine
    from oracle import weights
    cfg = LLAMA_CFGS[name]
    sd = _to(weights.llama(cfg, seed=5), dev)
    eng = engine.LlamaEngine(sd, cfg)
    assert eng.decode_plan is not None
    g = torch.Generator().manual_seed(B * 100 + T)
    x = torch.randn(B, T, cfg["hidden_size"], generator=g).to(bf16).to(dev)
    am = torch.ones(B, T + steps, dtype=torch.bool)
    if padded:
        am[1, 3:11] = False
    am = am.to(dev)
    caches = [eng.new_cache(B, T + steps + 2) for _ in range(2)]
    for c in caches:
        eng.forward(x.clone(), c, kv_mask=am)
    n0 = _launches()
    for s in range(steps):
        xs = torch.randn(B, 1, cfg["hidden_size"], generator=g).to(bf16).to(dev)
        outs = []
        for use, c in zip((True, False), caches):
            eng.use_decode_kernel = use
            before = _launches()
            outs.append(eng.forward(xs.clone(), c, kv_mask=am, want_router=True))
            if use:
                assert _launches() - before == 1, "the decode step must be ONE kernel launch"
        eng.use_decode_kernel = True
        torch.cuda.synchronize()
        a, b = outs
        if a["gate_logits"] is not None:
            _close(a["gate_logits"], b["gate_logits"], 2e-2, f"step {s} router logits")
            same = (a["gate_logits"].argmax(-1) == b["gate_logits"].argmax(-1)).all(0)  # per sequence, all layers
        else:
            same = torch.ones(B, dtype=torch.bool, device=dev)
        assert same.float().mean() >= 0.5
        _close(a["last_hidden_state"][same], b["last_hidden_state"][same], 2e-2, f"step {s} hidden")
        k0, k1 = caches[0].k[:, :, :, T + s], caches[1].k[:, :, :, T + s]
        v0, v1 = caches[0].v[:, :, :, T + s], caches[1].v[:, :, :, T + s]
        # layer 0 sees identical inputs: same RoPE roundings; the RMSNorm statistic is reduced in a different (fixed)
        # order, so a value may land on the other side of a bf16 rounding boundary once in a while
        _close(k0[0], k1[0], 2 ** -7, f"step {s} layer-0 k row")
        _close(v0[0], v1[0], 2 ** -7, f"step {s} layer-0 v row")
        _close(k0[:, same], k1[:, same], 2e-2, f"step {s} k rows")
        _close(v0[:, same], v1[:, same], 2e-2, f"step {s} v rows")
        # keep both caches on the same trajectory for the next step
        caches[1].k.copy_(caches[0].k)
        caches[1].v.copy_(caches[0].v)
    assert caches[0].len == T + steps and _launches() > n0


def _launches():
    from medplib_b200 import _lib
    return _lib.load().mpl_launch_count()
