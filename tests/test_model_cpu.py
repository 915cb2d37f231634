"""CPU: host-side logic of the drop-in classes — module tree / parameter names (the reference's naming invariants,
SURVEY.md §8b), MoE construction, the splice plan and seg-token mask against the golden vectors, and the loud failure
of the compute path without a GPU."""
import os
import sys
import types

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import inputs as gi  # noqa: E402

CLIP_CFG = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, image_size=28,
                patch_size=14, layer_norm_eps=1e-5)


@pytest.fixture(scope="module")
def model():
    from medplib_b200.model import MedPLIBForCausalLM, MedPLIBMoELlamaConfig
    cfg = MedPLIBMoELlamaConfig(hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=2,
                                num_key_value_heads=2, vocab_size=120, rms_norm_eps=1e-5, max_position_embeddings=128,
                                mm_vision_select_layer=-2, mm_projector_type="mlp2x_gelu", max_sample_point=512)
    cfg.clip_config = CLIP_CFG
    cfg.sam_config = dict(image_size=256, embed_dim=64, depth=2, num_heads=1)
    return MedPLIBForCausalLM(cfg, seg_token_idx=42, num_experts=[2], top_k_experts=1, capacity_factor=1.5,
                              eval_capacity_factor=2.0, min_capacity=0, use_residual=False, router_aux_loss_coef=0.01,
                              moe_layers_idx=None, ep_size=1, train_mask_decoder=True, out_dim=256, ce_loss_weight=1.0,
                              dice_loss_weight=0.5, bce_loss_weight=2.0, iou_loss_weight=2.0, focal_loss_weight=2.0,
                              mm_token_compress=True, icl_mask_encoder=True)


def test_parameter_names_follow_the_reference(model):
    names = set(n for n, _ in model.named_parameters())
    for n in ["model.embed_tokens.weight", "lm_head.weight", "model.norm.weight",
              "model.layers.0.self_attn.q_proj.weight", "model.layers.1.mlp.down_proj.weight",
              "model.layers.0.input_layernorm.weight", "model.layers.0.post_attention_layernorm.weight",
              "model.mm_projector.0.weight", "model.mm_projector.2.bias", "model.region_fea_adapter.weight",
              "model.mm_token_compressor.norm.weight", "model.mm_token_compressor.proj.weight",
              "model.mask_encoder.encoder.6.weight", "model.mask_encoder.proj.weight", "model.mask_encoder.norm.bias",
              "model.text_hidden_fcs.0.0.weight", "model.text_hidden_fcs.0.2.bias",
              "model.vision_tower.vision_tower.vision_model.embeddings.class_embedding",
              "model.vision_tower.vision_tower.vision_model.pre_layrnorm.weight",
              "model.vision_tower.vision_tower.vision_model.encoder.layers.1.self_attn.out_proj.bias",
              "model.visual_model.image_encoder.pos_embed", "model.visual_model.image_encoder.blocks.0.attn.rel_pos_h",
              "model.visual_model.image_encoder.blocks.1.Adapter.spatial.2.weight",
              "model.visual_model.image_encoder.neck.3.bias", "model.visual_model.prompt_encoder.no_mask_embed.weight",
              "model.visual_model.mask_decoder.transformer.layers.1.cross_attn_image_to_token.out_proj.weight",
              "model.visual_model.mask_decoder.output_hypernetworks_mlps.3.layers.2.bias",
              "model.visual_model.mask_decoder.iou_prediction_head.layers.0.weight"]:
        assert n in names, n
    assert "model.visual_model.prompt_encoder.pe_layer.positional_encoding_gaussian_matrix" in model.state_dict()
    # frozen / trainable split of initialize_bird_modules (MedPLIB.py:141-164)
    p = dict(model.named_parameters())
    assert not p["model.visual_model.image_encoder.pos_embed"].requires_grad
    assert p["model.visual_model.mask_decoder.iou_token.weight"].requires_grad
    assert p["model.text_hidden_fcs.0.0.weight"].requires_grad
    assert not p["model.vision_tower.vision_tower.vision_model.pre_layrnorm.weight"].requires_grad


def test_initialize_moe_modules_and_naming_invariants(model):
    args = types.SimpleNamespace(expert_pretrained_path="", moe_enable=True, moe_mode="dense", moe_layers_idx=None,
                                 ep_size=1, top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0,
                                 min_capacity=0, use_residual=False, router_aux_loss_coef=0.0, num_experts=[2])
    dense = model.model.layers[0].mlp.gate_proj.weight.detach().clone()
    model.initialize_moe_modules(args)
    names = dict(model.named_parameters())
    for l in (0, 1):
        wg = names[f"model.layers.{l}.mlp.deepspeed_moe.gate.wg.weight"]
        assert wg.dtype == torch.float32 and wg.shape == (2, 64)
        for e in (0, 1):
            for n in ("gate_proj", "up_proj", "down_proj"):
                w = names[f"model.layers.{l}.mlp.deepspeed_moe.experts.deepspeed_experts.{e}.{n}.weight"]
                assert getattr(w, "allreduce") is False  # DeepSpeed's expert-parameter tag
    # experts start as copies of the dense MLP (deepspeed Experts deep-copies the expert module)
    assert torch.equal(names["model.layers.0.mlp.deepspeed_moe.experts.deepspeed_experts.1.gate_proj.weight"], dense)
    assert model.config.moe["moe_layers_idx"] == [0, 1] and model.config.moe["num_experts"] == [2, 2]
    # LoRA target discovery of train_ds_medplib.py:265-291 finds the expert / attention linears by name
    targets = {n.split(".")[-1] for n, m in model.named_modules() if isinstance(m, torch.nn.Linear)
               and not any(x in n for x in ("visual_model", "vision_tower", "mm_projector", "text_hidden_fcs"))}
    assert {"q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj", "wg"} <= targets
    # the gate stays fp32 through .to(bf16) (DeepSpeed TopKGate keeps wg in fp32)
    model.to(torch.bfloat16)
    assert dict(model.named_parameters())["model.layers.0.mlp.deepspeed_moe.gate.wg.weight"].dtype == torch.float32
    assert model.lm_head.weight.dtype == torch.bfloat16
    model.to(torch.float32)


def test_seg_token_mask_matches_reference_golden(model):
    g = torch.load(os.path.join(HERE, "golden", "heads.pt"), weights_only=False)
    d = gi.heads_inputs()
    assert torch.equal(model.build_seg_token_mask(d["ids"], image_token_len=5), g["seg_mask"])
    assert torch.equal(model.build_seg_token_mask(d["ids"], image_token_len=5, image_token_lengths=[[3], [2, 4]]),
                       g["seg_mask_lengths"])


def test_compute_path_fails_loudly_on_cpu(model):
    from medplib_b200 import _lib
    ids = torch.randint(3, 100, (1, 8))
    with pytest.raises(_lib.MplError):
        model.generate(input_ids=ids, images=None, max_new_tokens=2)
    with pytest.raises(_lib.MplError):
        model.get_visual_embs(torch.zeros(1, 3, 256, 256))


def test_lisa_accepts_both_mask_spellings():
    import inspect
    from medplib_b200.model import LISAForCausalLM
    sig = inspect.signature(LISAForCausalLM.model_forward)
    assert "attention_masks" in sig.parameters  # LISA.py:267


def test_evaluate_seg_bookkeeping_on_the_host(model, monkeypatch):
    """evaluate() (model/MedPLIB.py:574-680) with the device work scripted: the row handed to text_hidden_fcs is the
    hidden state in front of the FIRST <SEG> of the generated ids (what `hidden[seg_mask][:1]` selects in the reference),
    position -2 when there is none, and `inference_demo` returns no mask without one. The product does this on ONE host
    copy of the ids; the expectation here is computed the reference's way (boolean indexing with the seg-token mask)."""
    from medplib_b200.model.MedPLIB import GenerateOutput
    SEG = model.seg_token_idx
    n_img = (getattr(model.config, "mm_compressed_token_count", 256) if getattr(model.config, "mm_token_compress", False)
             else model.get_model().get_vision_tower().num_patches)
    picked = {}

    def run(ids_row, demo=False):
        ids = torch.tensor([ids_row])
        T = len(ids_row) - 1 + n_img  # one -200 sentinel expands to n_img rows
        hidden = torch.randn(1, T - 1, 8, generator=torch.Generator().manual_seed(len(ids_row)))
        monkeypatch.setattr(model, "generate", lambda **kw: GenerateOutput(sequences=ids, hidden_states=None,
                                                                           last_hidden_state=hidden, scores=None))
        monkeypatch.setattr(model, "_seg_embeddings", lambda rows: picked.__setitem__("rows", rows.clone()) or rows)
        monkeypatch.setattr(model, "get_visual_embs", lambda images: torch.zeros(1))
        monkeypatch.setattr(model, "_decode_masks", lambda pe, ie, rl, sizes: (["mask"], None))
        model.overlap_vision = False
        picked.clear()
        out = model.evaluate(torch.zeros(1), torch.zeros(1), ids, [(256, 256)], [torch.zeros(4, 4)], inference_demo=demo)
        mask = model.build_seg_token_mask(ids)[:, :hidden.shape[1]]
        return out, hidden, mask

    out, hidden, mask = run([1, 5, -200, 7, 8, SEG, 9, SEG, 2])          # two <SEG>: the first one counts
    assert out[1] == ["mask"] and torch.equal(picked["rows"], hidden[mask][:1])
    out, hidden, mask = run([1, -200, 7, SEG, 2], demo=True)
    assert out[1] == ["mask"] and torch.equal(picked["rows"], hidden[mask][:1])
    out, hidden, mask = run([1, 5, -200, 7, 8, 9, 2])                     # none: position -2 (MedPLIB.py:640-644)
    assert int(mask.sum()) == 0 and torch.equal(picked["rows"], hidden[:1, -2])
    out, hidden, mask = run([1, 5, -200, 7, 8, 9, 2], demo=True)
    assert out[1] == [] and "rows" not in picked
