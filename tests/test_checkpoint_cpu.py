"""CPU tests of the construction / checkpoint contract (SURVEY.md §8b "Construction", §8 f-2): HF save_pretrained ->
from_pretrained(test_only=True) like model/eval/vqa_infer.py:226-237, the stage-3 -> stage-4 flow of
train_ds_medplib.py:225-232,310 (dense checkpoint, then initialize_moe_modules copies the MLP into the experts), the
DeepSpeed MoE checkpoint layout read by the reference's params_bf16_to_f32.py, and LoRA folding."""
import os
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bf16 = torch.bfloat16
CLIP_CFG = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2, image_size=28,
                patch_size=14, layer_norm_eps=1e-5)
KW = dict(seg_token_idx=42, num_experts=[2], top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0,
          min_capacity=0, use_residual=False, router_aux_loss_coef=0.01, moe_layers_idx=None, ep_size=1,
          train_mask_decoder=True, out_dim=256, ce_loss_weight=1.0, dice_loss_weight=0.5, bce_loss_weight=2.0,
          iou_loss_weight=2.0, focal_loss_weight=2.0)
MOE_ARGS = dict(expert_pretrained_path="", moe_enable=True, moe_mode="dense", moe_layers_idx=None, ep_size=1,
                top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0, min_capacity=0, use_residual=False,
                router_aux_loss_coef=0.01, num_experts=[2])


def _cfg():
    from medplib_b200.model import MedPLIBMoELlamaConfig
    cfg = MedPLIBMoELlamaConfig(hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=2,
                                num_key_value_heads=2, vocab_size=120, rms_norm_eps=1e-5, max_position_embeddings=128,
                                mm_vision_select_layer=-2, mm_projector_type="mlp2x_gelu", max_sample_point=512)
    cfg.clip_config = CLIP_CFG
    cfg.sam_config = dict(image_size=256, embed_dim=64, depth=2, num_heads=1)
    return cfg


def _randomize(m, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)


def _dense():
    from medplib_b200.model import MedPLIBForCausalLM
    m = MedPLIBForCausalLM(_cfg(), **KW)
    _randomize(m)
    return m


def _moe():
    m = _dense()
    m.initialize_moe_modules(types.SimpleNamespace(**MOE_ARGS))
    _randomize(m, 1)
    return m


def _same(a, b, names=None):
    sa, sb = a.state_dict(), b.state_dict()
    assert set(sa) == set(sb)
    for k in names or sa:
        assert torch.allclose(sa[k].float(), sb[k].float(), atol=2e-3, rtol=1e-2), k


def test_from_pretrained_restores_every_tensor(tmp_path):
    """vqa_infer.py:226-237: from_pretrained(path, torch_dtype=bf16, low_cpu_mem_usage=True,
    ignore_mismatched_sizes=True, test_only=True, **args) builds the MoE layers from config.moe and loads all weights —
    none may be re-initialised after loading; the gate stays fp32."""
    from medplib_b200.model import MedPLIBForCausalLM
    m = _moe()
    m.save_pretrained(tmp_path)
    m2 = MedPLIBForCausalLM.from_pretrained(tmp_path, torch_dtype=bf16, low_cpu_mem_usage=True,
                                            ignore_mismatched_sizes=True, test_only=True, **KW)
    _same(m, m2)
    sd = m2.state_dict()
    assert sd["model.layers.0.mlp.deepspeed_moe.gate.wg.weight"].dtype == torch.float32
    assert sd["lm_head.weight"].dtype == bf16
    assert m2.config.moe["moe_layers_idx"] == [0, 1] and m2.config.moe["num_experts"] == [2, 2]


def test_dense_checkpoint_then_initialize_moe_modules(tmp_path):
    """train_ds_medplib.py:225-232,310: a dense (stage-3) checkpoint is loaded, then initialize_moe_modules turns the
    MLPs into MoE layers whose experts start as copies of the dense MLP (medplib_moe_llama.py:604-635)."""
    from medplib_b200.model import MedPLIBForCausalLM
    d = _dense()
    d.save_pretrained(tmp_path)
    m = MedPLIBForCausalLM.from_pretrained(tmp_path, torch_dtype=bf16, low_cpu_mem_usage=True,
                                           ignore_mismatched_sizes=True, **KW)
    src = d.state_dict()
    assert torch.allclose(m.state_dict()["model.layers.1.mlp.up_proj.weight"].float(),
                          src["model.layers.1.mlp.up_proj.weight"], atol=2e-3)
    m.initialize_moe_modules(types.SimpleNamespace(**MOE_ARGS))
    sd = m.state_dict()
    for l in range(2):
        for e in range(2):
            for n in ("gate_proj", "up_proj", "down_proj"):
                k = f"model.layers.{l}.mlp.deepspeed_moe.experts.deepspeed_experts.{e}.{n}.weight"
                assert torch.allclose(sd[k].float(), src[f"model.layers.{l}.mlp.{n}.weight"], atol=2e-3), k
        assert sd[f"model.layers.{l}.mlp.deepspeed_moe.gate.wg.weight"].shape == (2, 64)


def test_deepspeed_moe_layout_round_trip(tmp_path):
    """The layout params_bf16_to_f32.py:5-28 merges: the model under "module" plus one file per expert; wrappers'
    prefixes (DeepSpeed `module.`, peft `base_model.model.`) are stripped on load."""
    from medplib_b200 import checkpoint as ck
    m = _moe()
    n = ck.save_deepspeed_layout(m, tmp_path, tag_prefix="base_model.model.")
    assert n == 4
    files = sorted(os.listdir(tmp_path))
    assert "mp_rank_00_model_states.pt" in files and "layer_1_expert_0_mp_rank_00_model_states.pt" in files
    merged = ck.merge_deepspeed_states(tmp_path)
    assert all(v.dtype == torch.float32 for v in merged.values() if torch.is_tensor(v) and v.is_floating_point())
    m2 = _moe()
    _randomize(m2, 5)
    missing, unexpected = ck.load_into(m2, merged)
    assert not missing and not unexpected
    _same(m, m2)
    # the same tensor in two files is an error, like the reference's script
    torch.save({"module": {"base_model.model.lm_head.weight": torch.zeros(1)}}, os.path.join(tmp_path, "zz_model_states.pt"))
    with pytest.raises(ValueError):
        ck.merge_deepspeed_states(tmp_path)


def test_lora_checkpoint_fold_and_keep(tmp_path):
    """A peft-layout checkpoint (`base_layer.weight`, `lora_A/B.default.weight`): folded into the base weights
    (merge_and_unload, merge_lora_weights_and_save_hf_model.py:183) or loaded into attached adapters."""
    from medplib_b200 import checkpoint as ck
    from medplib_b200 import train
    m = _moe()
    train.attach_lora(m, r=4, lora_alpha=8, target_modules="q_proj,v_proj,gate_proj")
    _randomize(m, 2)
    sd = {}
    for k, v in m.state_dict().items():  # re-key like peft: base weight of an adapted Linear -> base_layer.weight
        mod = k.rsplit(".", 1)[0]
        if k.endswith(".weight") and (mod + ".lora_A.default.weight") in m.state_dict():
            k = mod + ".base_layer.weight"
        sd["base_model.model." + k] = v.detach().clone()
    # keep: into a model that has the adapters
    m_keep = _moe()
    train.attach_lora(m_keep, r=4, lora_alpha=8, target_modules="q_proj,v_proj,gate_proj")
    missing, unexpected = ck.load_into(m_keep, sd, lora="keep")
    assert not missing and not unexpected
    _same(m, m_keep)
    # fold: into a plain model
    m_fold = _moe()
    missing, unexpected = ck.load_into(m_fold, sd, lora="fold", scaling=2.0)
    assert not missing and not unexpected
    ref = train.merge_lora(m)
    _same(ref, m_fold)
