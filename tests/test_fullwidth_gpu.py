"""GPU end-to-end parity at the 7B WIDTHS of the BASELINE configs (hidden 4096, FFN 11008, 32 heads x 128, vocab 32267,
CLIP-L width 1024 / 16 heads at 336 px, SAM-Med2D ViT-B width 768 / 12 heads, 2 experts top-1) with fewer LAYERS
(2 decoder layers, 4 CLIP layers, 3 SAM blocks incl. a global-attention one) so the CPU oracle finishes in seconds:

  * evaluate()  (configs[1]): prompt T = 40 + 575 = 615, 8 generated tokens with <SEG> at new token 4
  * MedPLIB-ICL separate mode (configs[4]): 3 (image, mask) exemplars + query, 576 -> 256 compression, 64-token mask
    encoder, T = 1299, model_forward(inference=True)

against BOTH oracles on the same weights: the eager-bf16 one (the reference's own arithmetic: stated tolerance, the
test's gate) and the fp32 one (how far the bf16 path is from north_star's fp32 statement "text logits 1e-4, mask logits
1e-3": measured and logged, gated loosely). Bit-exact claims: greedy token ids wherever the bf16 oracle's top-2 margin
exceeds the logits' measured noise, mask indices (logit > logit(0.1), vqa_infer.py:565) wherever the oracle's logit is
farther from the threshold than the mask logits' measured noise. Measured errors: profiles/r02_parity_errors.md."""
import math

import pytest
import torch

from parity import close, close_rows

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
SEG = 32003
D, F, H, V = 4096, 11008, 32, 32267
CLIP_CFG = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=4, num_attention_heads=16, image_size=336,
                patch_size=14, layer_norm_eps=1e-5)


def build_full(dev, icl=False, layers=2):
    from medplib_b200.model import MedPLIBForCausalLM, MedPLIBMoELlamaConfig
    torch.manual_seed(0)
    cfg = MedPLIBMoELlamaConfig(hidden_size=D, intermediate_size=F, num_hidden_layers=layers, num_attention_heads=H,
                                num_key_value_heads=H, vocab_size=V, rms_norm_eps=1e-5, max_position_embeddings=4096,
                                mm_vision_select_layer=-2, mm_projector_type="mlp2x_gelu", max_sample_point=512)
    cfg.clip_config = CLIP_CFG
    cfg.sam_config = dict(image_size=256, embed_dim=768, depth=3, num_heads=12)
    cfg.moe = dict(num_experts=[2], top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0, min_capacity=0,
                   use_residual=False, router_aux_loss_coef=0.01, moe_layers_idx=None, moe_mode="dense", ep_size=1)
    kw = dict(mm_token_compress=True, mm_compressed_token_count=256, icl_mask_encoder=True,
              mask_encoder_token_count=64) if icl else {}
    m = MedPLIBForCausalLM(cfg, test_only=True, seg_token_idx=SEG, use_mm_start_end=True, **kw)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():  # experts differ, norms / biases / rel-pos non-trivial, router decisive
        for n, p in m.named_parameters():
            if "deepspeed_experts.1" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif "rel_pos" in n or "pos_embed" in n or n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            if "wg.weight" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            if "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
    m.config.mm_use_im_start_end = True
    m = m.to(bf16).to(dev).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sd.update({k: v.detach().cpu() for k, v in m.named_buffers()})
    ocfg = dict(clip=dict(hidden_size=1024, intermediate_size=4096, num_layers=4, num_heads=16, image_size=336,
                          patch_size=14),
                llama=dict(hidden_size=D, intermediate_size=F, num_layers=layers, num_heads=H, vocab_size=V,
                           rms_norm_eps=1e-5, max_position_embeddings=4096, rope_theta=1e4, moe=m.config.moe),
                sam=dict(num_heads=12), mm_use_im_start_end=True, mm_token_compress=icl,
                mm_compressed_token_count=256, mask_encoder_token_count=64)
    return m, sd, ocfg


def _f32(sd):
    return {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}


def _mask_indices_equal(got, want, noise, name):
    thr = math.log(0.1 / 0.9)
    want = want.float()
    far = (want - thr).abs() > noise
    assert far.float().mean() > 0.5, f"{name}: too few pixels away from the threshold to make the claim"
    assert torch.equal((got.float().cpu() > thr)[far], (want > thr)[far]), f"{name}: mask indices differ"
    return float(far.float().mean())


def test_evaluate_full_width(dev):
    from oracle import pipeline
    m, sd, ocfg = build_full(dev)
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(3, 31999, (1, 40), generator=g)
    ids[0, 2], ids[0, 3], ids[0, 4] = 32001, -200, 32002
    clip_img = torch.randn(1, 3, 336, 336, generator=g).to(bf16)
    sam_img = torch.randn(1, 3, 256, 256, generator=g).to(bf16)
    label = torch.zeros(336, 336)
    forced = {4: SEG}
    ref = pipeline.evaluate(sd, ocfg, clip_img, sam_img, ids, [(256, 256)], [(336, 336)], 8, SEG, forced_tokens=forced)
    assert ref["hidden"].shape[1] == 615 + 7
    ref32 = pipeline.evaluate(_f32(sd), ocfg, clip_img.float(), sam_img.float(), ids, [(256, 256)], [(336, 336)], 8, SEG,
                              forced_tokens={i: int(t) for i, t in enumerate(ref["output_ids"][0, 40:])})
    ref_new = ref["output_ids"][0, 40:]
    force_all = {i: int(t) for i, t in enumerate(ref_new)}
    # The router is a hard top-1 decision: a token whose two logits are a near-tie goes to the other expert under ANY
    # change of rounding (the reference's own bf16 and fp32 runs disagree on such tokens), its row then differs
    # wholesale and, through attention, perturbs every later token. So: (1) run the GPU path and record its routing;
    # (2) require it to equal the oracle's own routing wherever the oracle's margin exceeds the router's measured noise;
    # (3) compare every activation against the oracle evaluated UNDER THE GPU'S ROUTING (oracle.moe.forced_routing).
    from oracle import llama, moe as omoe
    emb_o, am_o, _ = pipeline.prefill_inputs(sd, ocfg, clip_img, ids, torch.ones_like(ids, dtype=torch.bool))
    pre = llama.model_forward(sd, ocfg["llama"], emb_o, am_o)
    seen = []
    hooks = [mod.register_forward_hook(lambda mod_, i, o: seen.append(o.detach().float().cpu()))
             for n, mod in m.named_modules() if "wg" in n and isinstance(mod, torch.nn.Linear)]
    gen = m.generate(input_ids=ids.to(dev), images=clip_img.to(dev), max_new_tokens=8, output_hidden_states=True,
                     return_dict_in_generate=True, output_scores=True, forced_tokens=force_all, eos_token_id=-1)
    for h in hooks:
        h.remove()
    assert torch.equal(gen.sequences.cpu(), ref["output_ids"])
    assert len(seen) == 2 * 8  # 2 MoE layers x (prefill + 7 KV-cached steps)
    flips = 0
    for l, want_lg in enumerate(pre["gate_logits"]):
        got_lg, want_lg = seen[l], want_lg.float()
        assert got_lg.shape == want_lg.shape == (615, 2)
        dlg = (got_lg - want_lg).abs().amax(-1)
        noise_l = dlg.quantile(0.99).item() if l == 0 else dlg.quantile(0.90).item()
        if l == 0:  # layer 0 sees the same inputs on both sides
            close(got_lg, want_lg, 2e-2, "layer 0 router logits vs bf16 oracle")
        margin = (want_lg[:, 0] - want_lg[:, 1]).abs()
        same = got_lg.argmax(-1) == want_lg.argmax(-1)
        flips += int((~same).sum())
        if l == 0:
            assert bool(same[margin > 4 * noise_l].all()), "layer 0: routing differs on a decisive token"
    assert flips <= 0.05 * 2 * 615, f"{flips} routing decisions differ from the oracle's"
    with omoe.forced_routing([s_.argmax(-1) for s_ in seen]):
        ref = pipeline.evaluate(sd, ocfg, clip_img, sam_img, ids, [(256, 256)], [(336, 336)], 8, SEG, forced_tokens=force_all)
    with omoe.forced_routing([s_.argmax(-1) for s_ in seen]):
        ref32 = pipeline.evaluate(_f32(sd), ocfg, clip_img.float(), sam_img.float(), ids, [(256, 256)], [(336, 336)], 8,
                                  SEG, forced_tokens=force_all)
    # Stated tolerance, at 7B width and under identical routing: the bf16 path deviates from the reference's eager-bf16
    # arithmetic by no more than that arithmetic's OWN rounding noise, measured as the gap between the bf16 and the fp32
    # oracle on the same inputs (softmax over bf16-rounded q.k scores amplifies one-ulp differences of q and k: with these
    # random weights the reference's bf16 run sits ~15 % of max|hidden| away from its fp32 run on the worst row).
    own_gap = close(ref["hidden"], ref32["hidden"], 1.0,
                    "bf16 ORACLE vs fp32 oracle: hidden states (the reference's own bf16 rounding noise)")
    dev_gap = close(gen.last_hidden_state, ref["hidden"], 0.25,
                    f"hidden states vs bf16 oracle, same routing ({flips} of {2 * 615} prefill decisions were near-ties "
                    "that fell the other way)")
    assert dev_gap <= 1.25 * own_gap + 2.0 ** -8, (dev_gap, own_gap)
    noise = 0.0
    for s, (got, want) in enumerate(zip(gen.scores, ref["step_logits"])):
        noise = max(noise, close(got, want, 3e-2, f"step {s} logits vs bf16 oracle") * want.abs().max().item())
    for s, (got, want) in enumerate(zip(gen.scores, ref["step_logits"])):
        top2 = want[0].float().topk(2).values
        if s not in forced and (top2[0] - top2[1]) > 2 * noise:
            assert int(got[0].argmax()) == int(want[0].argmax()), f"argmax at step {s}"
    # distance to the fp32 statement of the same path (north_star: 1e-4): measured, logged, loosely gated
    close(gen.last_hidden_state, ref32["hidden"], 0.3, "hidden states vs fp32 oracle (same routing)")
    for s, (got, want) in enumerate(zip(gen.scores, ref32["step_logits"])):
        close(got, want, 2.5e-2, f"step {s} logits vs fp32 oracle")
    out_ids, masks = m.evaluate(clip_img.to(dev), sam_img.to(dev), ids.to(dev), [(256, 256)], [label],
                                max_new_tokens=8, forced_tokens=force_all)
    assert torch.equal(out_ids.cpu(), ref["output_ids"])
    want = ref["pred_masks"][0]
    rel = close(masks[0], want, 2.5e-2, "mask logits vs bf16 oracle")
    close(masks[0], ref32["pred_masks"][0], 2.5e-2, "mask logits vs fp32 oracle")
    close(want, ref32["pred_masks"][0], 8e-2, "bf16 ORACLE vs fp32 oracle: mask logits (the reference's own gap)")
    _mask_indices_equal(masks[0], want, 2 * rel * want.float().abs().max().item() + 1e-3, "evaluate")


def test_icl_full_width(dev):
    from oracle import pipeline
    m, sd, ocfg = build_full(dev, icl=True)
    g = torch.Generator().manual_seed(4)
    n_text, n_img, n_mask = 90, 4, 3
    types_ = [["image", "mask"] * n_mask + ["image"]]
    lengths = [[256, 64] * n_mask + [256]]
    ids = torch.randint(3, 31999, (1, n_text), generator=g)
    for k in range(n_img + n_mask):
        ids[0, 4 + 8 * k] = -200
    ids[0, n_text - 6] = SEG
    clip_imgs = [torch.randn(n_img, 3, 336, 336, generator=g).to(bf16)]
    mask_imgs = [(torch.rand(n_mask, 1, 336, 336, generator=g) < 0.2).to(bf16)]
    sam_img = torch.randn(1, 3, 256, 256, generator=g).to(bf16)
    label = torch.zeros(336, 336)
    ref = pipeline.grounding_forward_icl(sd, ocfg, clip_imgs, mask_imgs, types_, lengths, sam_img, ids, [(256, 256)],
                                         [(336, 336)], SEG)
    T = n_text - (n_img + n_mask) + n_img * 256 + n_mask * 64
    assert T == 1299 and ref["hidden"].shape[1] == T
    ref32 = pipeline.grounding_forward_icl(_f32(sd), ocfg, [c.float() for c in clip_imgs], [x.float() for x in mask_imgs],
                                           types_, lengths, sam_img.float(), ids, [(256, 256)], [(336, 336)], SEG)
    am = torch.ones_like(ids, dtype=torch.bool).to(dev)
    out = m(images=sam_img.to(dev), images_clip=[c.to(dev) for c in clip_imgs], input_ids=ids.to(dev), region_masks=None,
            labels=None, attention_mask=am, offset=None, masks_list=[label], label_list=[label],
            resize_list=[(256, 256)], inference=True, mask_images=[x.to(dev) for x in mask_imgs],
            image_token_types=types_, image_token_lengths=lengths, icl_image_counts=[n_img])
    _, _, _, emb, _ = m.prepare_inputs_labels_for_multimodal(
        ids.to(dev), am, None, None, [c.to(dev) for c in clip_imgs], None, None,
        mask_images=[x.to(dev) for x in mask_imgs], image_token_types=types_)
    assert emb.shape[1] == T
    close(emb, ref["inputs_embeds"], 1.4e-2, "ICL inputs_embeds (CLIP x4 + compressor + mask encoder + splice) vs bf16 oracle")
    close(emb, ref32["inputs_embeds"], 1.3e-2, "ICL inputs_embeds vs fp32 oracle")
    want = ref["pred_masks"][0]
    rel = close(out["pred_masks"][0], want, 2.5e-2, "ICL mask logits vs bf16 oracle")
    close(out["pred_masks"][0], ref32["pred_masks"][0], 2.5e-2, "ICL mask logits vs fp32 oracle")
    close(want, ref32["pred_masks"][0], 8e-2, "bf16 ORACLE vs fp32 oracle: ICL mask logits (the reference's own gap)")
    _mask_indices_equal(out["pred_masks"][0], want, 2 * rel * want.float().abs().max().item() + 1e-3, "ICL")
