"""GPU end-to-end parity of MedPLIBForCausalLM (evaluate / model_forward(inference=True) / generate) against the CPU
oracle pipeline on a small random-init model with the reference's architecture (MoE top-1, 2 experts, SAM adapters).

Stated tolerances (bf16 path vs bf16 oracle; measured values in profiles/r02_parity_errors.md): hidden states and mask
logits within 4.5e-2 / 4e-2 * max|ref|, step logits within 3.5e-2 (measured 2.5e-2 / 2.4e-2 / 2.0e-2); greedy token ids
bit-exact wherever the oracle's top-2 logit margin exceeds that noise; mask indices (logit > logit(0.1), the
reference's `sigmoid(pred) > 0.1`, vqa_infer.py:565) bit-exact wherever the oracle's logit is farther than the noise
from the threshold."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
SEG = 299
CLIP_CFG = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=56,
                patch_size=14, layer_norm_eps=1e-5)


def build(dev, compress=False, moe_layers=None):
    from medplib_b200.model import MedPLIBForCausalLM, MedPLIBMoELlamaConfig
    torch.manual_seed(0)
    cfg = MedPLIBMoELlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                                num_key_value_heads=2, vocab_size=300, rms_norm_eps=1e-5, max_position_embeddings=512,
                                mm_vision_select_layer=-2, mm_projector_type="mlp2x_gelu", max_sample_point=512,
                                initializer_range=0.06)
    cfg.clip_config = CLIP_CFG
    cfg.sam_config = dict(image_size=256, embed_dim=128, depth=3, num_heads=2)
    cfg.moe = dict(num_experts=[2], top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0, min_capacity=0,
                   use_residual=False, router_aux_loss_coef=0.01, moe_layers_idx=moe_layers, moe_mode="dense", ep_size=1)
    m = MedPLIBForCausalLM(cfg, test_only=True, seg_token_idx=SEG, mm_token_compress=compress, use_mm_start_end=True)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():  # make experts differ, norms / biases / rel-pos non-trivial, router decisive
        for n, p in m.named_parameters():
            if "deepspeed_experts.1" in n or "rel_pos" in n or "pos_embed" in n or n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.06)
            if "wg.weight" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            if "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
    m.config.mm_use_im_start_end = True
    m = m.to(bf16).to(dev).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sd.update({k: v.detach().cpu() for k, v in m.named_buffers()})
    ocfg = dict(clip=dict(hidden_size=128, intermediate_size=256, num_layers=3, num_heads=2, image_size=56,
                          patch_size=14),
                llama=dict(hidden_size=256, intermediate_size=512, num_layers=2, num_heads=2, vocab_size=300,
                           rms_norm_eps=1e-5, max_position_embeddings=512, rope_theta=1e4, moe=m.config.moe),
                sam=dict(num_heads=2), mm_use_im_start_end=True, mm_token_compress=compress)
    return m, sd, ocfg


def inputs(n_text=12, seg_in_prompt=False):
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(3, 290, (1, n_text), generator=g)
    ids[0, 2] = -200
    if seg_in_prompt:
        ids[0, 8] = SEG
    clip_img = torch.randn(1, 3, 56, 56, generator=g).to(bf16)
    sam_img = torch.randn(1, 3, 256, 256, generator=g).to(bf16)
    return ids, clip_img, sam_img


def _check(got, ref, rtol, name):
    from parity import close
    return close(got, ref, rtol, name)


def test_evaluate_matches_oracle(dev):
    from oracle import pipeline
    m, sd, ocfg = build(dev)
    ids, clip_img, sam_img = inputs()
    label = torch.zeros(70, 90)
    forced = {3: SEG}
    ref = pipeline.evaluate(sd, ocfg, clip_img, sam_img, ids, [(256, 256)], [tuple(label.shape)], 6, SEG,
                            forced_tokens=forced)
    ref_new = ref["output_ids"][0, ids.shape[1]:]
    # feed the oracle's tokens so both sides follow the same trajectory; compare per-step logits and argmax
    force_all = {i: int(t) for i, t in enumerate(ref_new)}
    gen = m.generate(input_ids=ids.to(dev), images=clip_img.to(dev), max_new_tokens=6, output_hidden_states=True,
                     return_dict_in_generate=True, output_scores=True, forced_tokens=force_all, eos_token_id=-1)
    assert torch.equal(gen.sequences.cpu(), ref["output_ids"])
    _check(gen.last_hidden_state, ref["hidden"], 4.5e-2, "hidden states")
    for s, (got, want) in enumerate(zip(gen.scores, ref["step_logits"])):
        _check(got, want, 3.5e-2, f"step {s} logits")
        top2 = want[0].topk(2).values
        if s not in forced and (top2[0] - top2[1]) > 4e-2 * want.abs().max():  # 2 x the measured logit noise
            assert int(got[0].argmax()) == int(want[0].argmax()), f"argmax at step {s}"
    out_ids, masks = m.evaluate(clip_img.to(dev), sam_img.to(dev), ids.to(dev), [(256, 256)], [label],
                                max_new_tokens=6, forced_tokens=force_all)
    assert torch.equal(out_ids.cpu(), ref["output_ids"])
    assert masks[0].shape == (1, 70, 90)
    _check(masks[0], ref["pred_masks"][0], 4e-2, "mask logits")
    thr = math.log(0.1 / 0.9)
    want = ref["pred_masks"][0].float()
    far = (want - thr).abs() > 4e-2 * want.abs().max()
    assert far.float().mean() > 0.5
    assert torch.equal((masks[0].float().cpu() > thr)[far], (want > thr)[far]), "mask indices"


def test_grounding_forward_matches_oracle(dev):
    from oracle import pipeline
    m, sd, ocfg = build(dev)
    ids, clip_img, sam_img = inputs(seg_in_prompt=True)
    label = torch.zeros(336, 336)
    ref = pipeline.grounding_forward(sd, ocfg, clip_img, sam_img, ids, [(256, 256)], [(336, 336)], SEG)
    out = m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), region_masks=None,
            labels=None, attention_mask=torch.ones_like(ids, dtype=torch.bool).to(dev), offset=None,
            masks_list=[label], label_list=[label], resize_list=[(256, 256)], inference=True)
    assert set(out) == {"pred_masks", "gt_masks"}
    _check(out["pred_masks"][0], ref["pred_masks"][0], 3e-2, "mask logits")


def test_gate_hooks_and_lm_forward(dev):
    """vqa_infer.py:157-165: forward hooks on the `wg` Linears observe the router logits."""
    m, sd, ocfg = build(dev)
    ids, clip_img, _ = inputs()
    seen = []
    for n, mod in m.named_modules():
        if "wg" in n and isinstance(mod, torch.nn.Linear):
            mod.register_forward_hook(lambda mod_, i, o: seen.append(o.detach().cpu()))
    out = m(input_ids=ids.to(dev), images=clip_img.to(dev), past_key_values=None, use_cache=True,
            attention_mask=torch.ones_like(ids, dtype=torch.bool).to(dev))
    T = ids.shape[1] - 1 + 16
    assert out.logits.shape == (1, T, 300) and out.logits.dtype == torch.float32
    assert len(seen) == 2 and seen[0].shape == (T, 2)
    assert out.past_key_values.len == T
    nxt = out.logits[:, -1].argmax(-1, keepdim=True)
    out2 = m(input_ids=nxt, images=clip_img.to(dev), past_key_values=out.past_key_values, use_cache=True,
             attention_mask=torch.ones((1, T + 1), dtype=torch.bool, device=dev))
    assert out2.logits.shape == (1, 1, 300) and out2.past_key_values.len == T + 1


def test_gate_hooks_sparse_moe_layout(dev):
    """MoE in layer 1 only (--moe_mode second_half / explicit moe_layers_idx): the single `wg` hook must observe layer
    1's router logits (the engine's buffer is indexed by transformer layer), for the prefill and for a decode step
    through the persistent kernel; moe_loss_list has one entry per MoE layer (medplib_moe_llama.py:265-283)."""
    from oracle import llama, pipeline
    m, sd, ocfg = build(dev, moe_layers=[1])
    ids, clip_img, _ = inputs()
    hooked = [n for n, mod in m.named_modules() if "wg" in n and isinstance(mod, torch.nn.Linear)]
    assert hooked == ["model.layers.1.mlp.deepspeed_moe.gate.wg"]
    seen = []
    m.get_submodule(hooked[0]).register_forward_hook(lambda mod_, i, o: seen.append(o.detach().float().cpu()))
    am = torch.ones_like(ids, dtype=torch.bool)
    out = m(input_ids=ids.to(dev), images=clip_img.to(dev), past_key_values=None, use_cache=True,
            attention_mask=am.to(dev))
    emb, am2, _ = pipeline.prefill_inputs(sd, ocfg, clip_img, ids, am)
    ref = llama.model_forward(sd, ocfg["llama"], emb, am2)
    assert len(ref["gate_logits"]) == 1 and len(seen) == 1 and len(out.moe_loss_list) == 1
    want = ref["gate_logits"][0].float()
    assert seen[0].shape == want.shape
    assert (seen[0] - want).abs().max() <= 4e-2 * want.abs().max()
    assert seen[0].abs().max() > 0.1 * want.abs().max()
    T = out.past_key_values.len
    nxt = out.logits[:, -1].argmax(-1, keepdim=True)
    m(input_ids=nxt, images=clip_img.to(dev), past_key_values=out.past_key_values, use_cache=True,
      attention_mask=torch.ones((1, T + 1), dtype=torch.bool, device=dev))
    assert len(seen) == 2 and seen[1].shape == (1, 2) and seen[1].abs().max() > 0


def test_cpu_tensors_fail_loudly(dev):
    from medplib_b200 import ops, _lib
    with pytest.raises(_lib.MplError):
        ops.linear(torch.zeros(4, 8, dtype=bf16), torch.zeros(8, 8, dtype=bf16))


def build_icl(dev):
    """BASELINE configs[4] at toy size: ICL separate mode, token compressor (16 -> 8 tokens per image) and the
    MaskTokenEncoder (4 tokens per exemplar mask)."""
    from medplib_b200.model import MedPLIBForCausalLM, MedPLIBMoELlamaConfig
    torch.manual_seed(0)
    cfg = MedPLIBMoELlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                                num_key_value_heads=2, vocab_size=300, rms_norm_eps=1e-5, max_position_embeddings=512,
                                mm_vision_select_layer=-2, mm_projector_type="mlp2x_gelu", max_sample_point=512,
                                initializer_range=0.06)
    cfg.clip_config = CLIP_CFG
    cfg.sam_config = dict(image_size=256, embed_dim=128, depth=3, num_heads=2)
    cfg.moe = dict(num_experts=[2], top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0, min_capacity=0,
                   use_residual=False, router_aux_loss_coef=0.01, moe_layers_idx=None, moe_mode="dense", ep_size=1)
    m = MedPLIBForCausalLM(cfg, test_only=True, seg_token_idx=SEG, mm_token_compress=True, mm_compressed_token_count=8,
                           icl_mask_encoder=True, mask_encoder_token_count=4, use_mm_start_end=True)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "deepspeed_experts.1" in n or "rel_pos" in n or "pos_embed" in n or n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.06)
            if "wg.weight" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            if "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
    m.config.mm_use_im_start_end = True
    m = m.to(bf16).to(dev).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sd.update({k: v.detach().cpu() for k, v in m.named_buffers()})
    ocfg = dict(clip=dict(hidden_size=128, intermediate_size=256, num_layers=3, num_heads=2, image_size=56,
                          patch_size=14),
                llama=dict(hidden_size=256, intermediate_size=512, num_layers=2, num_heads=2, vocab_size=300,
                           rms_norm_eps=1e-5, max_position_embeddings=512, rope_theta=1e4, moe=m.config.moe),
                sam=dict(num_heads=2), mm_use_im_start_end=True, mm_token_compress=True, mm_compressed_token_count=8,
                mask_encoder_token_count=4)
    return m, sd, ocfg


def test_icl_separate_mode_matches_oracle(dev):
    """MedPLIB-ICL separate mode (BASELINE configs[4]): two (image, mask) exemplars + the query image, compressed image
    tokens, mask-encoder tokens, [SEG] in the prompt, single-pass model_forward(inference=True)."""
    from oracle import pipeline
    m, sd, ocfg = build_icl(dev)
    g = torch.Generator().manual_seed(4)
    types_ = [["image", "mask", "image", "mask", "image"]]
    lengths = [[8, 4, 8, 4, 8]]
    ids = torch.randint(3, 290, (1, 26), generator=g)
    for k, pos in enumerate((2, 6, 10, 14, 18)):  # IMAGE sentinel followed by its <im_end> slot
        ids[0, pos] = -200
    ids[0, 23] = SEG
    clip_imgs = [torch.randn(3, 3, 56, 56, generator=g).to(bf16)]
    mask_imgs = [(torch.rand(2, 1, 56, 56, generator=g) > 0.8).to(bf16)]
    sam_img = torch.randn(1, 3, 256, 256, generator=g).to(bf16)
    label = torch.zeros(70, 90)
    ref = pipeline.grounding_forward_icl(sd, ocfg, clip_imgs, mask_imgs, types_, lengths, sam_img, ids, [(256, 256)],
                                         [(70, 90)], SEG)
    T_ref = ref["hidden"].shape[1]
    assert T_ref == 26 - 5 + 3 * 8 + 2 * 4
    out = m(images=sam_img.to(dev), images_clip=[c.to(dev) for c in clip_imgs], input_ids=ids.to(dev), region_masks=None,
            labels=None, attention_mask=torch.ones_like(ids, dtype=torch.bool).to(dev), offset=None,
            masks_list=[label], label_list=[label], resize_list=[(256, 256)], inference=True,
            mask_images=[x.to(dev) for x in mask_imgs], image_token_types=types_, image_token_lengths=lengths,
            icl_image_counts=[3])
    assert out["pred_masks"][0].shape == (1, 70, 90)
    _check(out["pred_masks"][0], ref["pred_masks"][0], 3.5e-2, "ICL mask logits")
    # the spliced prompt itself (compressor + mask encoder + sentinel order) against the oracle's inputs_embeds
    _, _, _, emb, _ = m.prepare_inputs_labels_for_multimodal(
        ids.to(dev), torch.ones_like(ids, dtype=torch.bool).to(dev), None, None, [c.to(dev) for c in clip_imgs], None,
        None, mask_images=[x.to(dev) for x in mask_imgs], image_token_types=types_)
    _check(emb, ref["inputs_embeds"], 1.6e-2, "ICL inputs_embeds")


@pytest.mark.parametrize("case", [0, 1])
def test_splice_matches_reference_golden(dev, case):
    """MedPLIBForCausalLM.prepare_inputs_labels_for_multimodal (host plan + one row-gather kernel) against the
    REFERENCE's own function (tests/golden/splice.pt, generated by tests/golden/make_golden.py): standard + region
    slot, plain, and ICL separate mode with ragged samples. Labels and attention masks bit-exact, embeddings to bf16."""
    import os
    import sys
    from medplib_b200.model import MedPLIBForCausalLM, MedPLIBMoELlamaConfig
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gdir)
    import inputs as gi
    g = torch.load(os.path.join(gdir, "splice.pt"), weights_only=False)["cases"][case]
    use_se = g["use_se"]
    cfg = MedPLIBMoELlamaConfig(hidden_size=16, intermediate_size=32, num_hidden_layers=1, num_attention_heads=1,
                                num_key_value_heads=1, vocab_size=50, rms_norm_eps=1e-5, max_position_embeddings=64,
                                mm_vision_select_layer=-2, mm_projector_type="mlp2x_gelu", max_sample_point=512)
    cfg.clip_config = CLIP_CFG
    cfg.sam_config = dict(image_size=256, embed_dim=64, depth=1, num_heads=1)
    cfg.moe = dict(num_experts=[2], top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0, min_capacity=0,
                   use_residual=False, router_aux_loss_coef=0.01, moe_layers_idx=None, moe_mode="dense", ep_size=1)
    m = MedPLIBForCausalLM(cfg, test_only=True, seg_token_idx=42, use_mm_start_end=use_se)
    m.config.mm_use_im_start_end = use_se
    m = m.to(bf16).to(dev).eval()

    def fake_images(images, region_flag=False, region_geo_sampler=False):
        x = images.to(bf16)
        return x, x, ((2 * images).to(bf16) if region_flag else None)

    m.encode_images = fake_images
    m.encode_masks = lambda masks: masks.to(bf16)

    def run(d, feats, region_masks=None, valid=None, **kw):
        with torch.no_grad():
            m.model.embed_tokens.weight.copy_(d["embed"].to(bf16))
        rm = [[x.to(dev) for x in r] for r in region_masks] if region_masks is not None else None
        f = [x.to(dev) for x in feats] if isinstance(feats, list) else feats.to(dev)
        _, am, _, emb, lab = m.prepare_inputs_labels_for_multimodal(
            d["ids"].to(dev), d["am"].to(dev), None, d["labels"].to(dev), f, rm, valid, **kw)
        return emb, lab, am

    d = gi.splice_inputs(use_se)
    emb, lab, am = run(d, d["feats_r"], d["region_masks"], d["valid"])
    assert torch.equal(lab.cpu(), g["lab1"]) and torch.equal(am.cpu(), g["am1"])
    _check(emb, g["emb1"], 6e-3, "region splice")
    d2 = dict(d, ids=d["ids2"])
    emb, lab, am = run(d2, d["feats"])
    assert torch.equal(lab.cpu(), g["lab2"]) and torch.equal(am.cpu(), g["am2"])
    _check(emb, g["emb2"], 6e-3, "plain splice")
    di = gi.icl_splice_inputs(use_se)
    emb, lab, am = run(di, di["img"], mask_images=[x.to(dev) for x in di["msk"]], image_token_types=di["types"])
    assert torch.equal(lab.cpu(), g["lab3"]) and torch.equal(am.cpu(), g["am3"])
    _check(emb, g["emb3"], 6e-3, "ICL splice")
