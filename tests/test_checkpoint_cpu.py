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


REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "params_bf16_to_f32.py")), reason="needs the reference tree")


def _ref_loader():
    """The reference's own merge function (params_bf16_to_f32.py:5-28), imported from where it lies."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_params_bf16_to_f32", os.path.join(REF, "params_bf16_to_f32.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.load_model_parameters


@needs_ref
def test_reference_merge_script_reads_the_written_layout(tmp_path):
    """save_deepspeed_layout -> the REFERENCE's params_bf16_to_f32.load_model_parameters -> load_into: every tensor comes
    back. With adapters attached the keys are exactly the ones a peft-wrapped model has (what
    merge_lora_weights_and_save_hf_model_moe.py loads with strict=False: a key without the `base_model.model.` prefix
    or without `base_layer` would be dropped silently and the merge would emit an untrained model)."""
    from medplib_b200 import checkpoint as ck
    from medplib_b200 import train
    from medplib_b200.compat import peft_shim
    m = _moe()
    train.attach_lora(m, r=4, lora_alpha=8, target_modules="q_proj,v_proj,gate_proj,up_proj,down_proj")
    _randomize(m, 3)
    assert ck.save_deepspeed_layout(m, tmp_path) == 4
    merged = _ref_loader()(str(tmp_path), "cpu")
    assert all(v.dtype == torch.float32 for v in merged.values())
    # the key set of a peft-wrapped model of the same architecture (the stand-in reproduces peft's nesting)
    twin = _moe()
    wrapped = peft_shim.get_peft_model(twin, peft_shim.LoraConfig(r=4, lora_alpha=8, lora_dropout=0.0, target_modules=[
        n for n, mod in twin.named_modules() if isinstance(mod, torch.nn.Linear)
        and any(t in n for t in ("q_proj", "v_proj", "gate_proj", "up_proj", "down_proj"))
        and not any(x in n for x in ("visual_model", "vision_tower", "mm_projector"))]))
    assert set(merged) == set(wrapped.state_dict())
    assert "base_model.model.model.layers.0.self_attn.q_proj.base_layer.weight" in merged
    # ... and loads straight into that wrapped model, like the reference's merge script does (strict=False, nothing dropped)
    res = wrapped.load_state_dict({k: v.to(wrapped.state_dict()[k].dtype) for k, v in merged.items()}, strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    folded = wrapped.merge_and_unload()
    _same(train.merge_lora(m), folded)


@needs_ref
def test_expert_files_are_numbered_by_moe_layer(tmp_path):
    """DeepSpeed names expert files by moe_layer_id — a counter over MoE layers — not by transformer layer: with MoE in
    layer 1 only (--moe_mode second_half on 2 layers) the files are layer_0_expert_{0,1}_…; the keys inside keep the
    transformer layer index (medplib_moe_llama.py:617-635)."""
    from medplib_b200 import checkpoint as ck
    m = _dense()
    m.initialize_moe_modules(types.SimpleNamespace(**dict(MOE_ARGS, moe_mode="second_half")))
    _randomize(m, 4)
    assert ck.save_deepspeed_layout(m, tmp_path) == 2
    files = sorted(f for f in os.listdir(tmp_path) if "expert" in f)
    assert files == ["layer_0_expert_0_mp_rank_00_model_states.pt", "layer_0_expert_1_mp_rank_00_model_states.pt"]
    merged = _ref_loader()(str(tmp_path), "cpu")
    assert "model.layers.1.mlp.deepspeed_moe.experts.deepspeed_experts.1.up_proj.weight" in merged
    m2 = _dense()
    m2.initialize_moe_modules(types.SimpleNamespace(**dict(MOE_ARGS, moe_mode="second_half")))
    missing, unexpected = ck.load_into(m2, merged)
    assert not missing and not unexpected
    _same(m, m2)


def test_compat_engine_checkpoint_round_trip(tmp_path):
    """deepspeed stand-in: engine.save_checkpoint(dir) writes <dir>/<tag>/… + <dir>/latest like DeepSpeed
    (train_ds_medplib.py:693-698), engine.load_checkpoint(dir) reads `latest` and restores the peft-wrapped model and
    the step counters (auto-resume, train_ds_medplib.py:452-470)."""
    from medplib_b200.compat import deepspeed_shim as ds
    from medplib_b200.compat import peft_shim
    cfg = {"train_micro_batch_size_per_gpu": 2, "gradient_accumulation_steps": 2,
           "optimizer": {"type": "AdamW", "params": {"lr": 3e-4, "weight_decay": 0.0, "betas": (0.9, 0.95)}},
           "scheduler": {"type": "WarmupDecayLR", "params": {"total_num_steps": 100, "warmup_min_lr": 0, "warmup_max_lr": 3e-4,
                                                             "warmup_num_steps": 10, "warmup_type": "linear"}},
           "gradient_clipping": 1.0, "bf16": {"enabled": True}}

    def wrapped(seed):
        m = _moe()
        w = peft_shim.get_peft_model(m, peft_shim.LoraConfig(r=4, lora_alpha=8, target_modules=["q_proj", "v_proj"]))
        _randomize(w, seed)
        return w

    a = wrapped(6)
    eng, opt, loader, sched = ds.initialize(model=a, model_parameters=a.parameters(), config=cfg)
    assert opt is None and loader is None and sched is eng.lr_scheduler
    eng.global_steps, eng.micro_steps = 7, 14
    for _ in range(7):
        eng.lr_scheduler.step()
    eng.save_checkpoint(str(tmp_path))
    assert open(os.path.join(tmp_path, "latest")).read().strip() == "global_step7"
    b = wrapped(9)
    eng2, _, _, _ = ds.initialize(model=b, model_parameters=b.parameters(), config=cfg)
    path, client = eng2.load_checkpoint(str(tmp_path))
    assert path.endswith("global_step7") and client == {}
    assert eng2.global_steps == 7 and eng2.get_lr() == eng.get_lr()
    _same(a, b)
    assert eng2.load_checkpoint(os.path.join(tmp_path, "nothing_here")) == (None, None)
    # linear warm-up then linear decay (deepspeed/runtime/lr_schedules.py::WarmupDecayLR)
    s = ds.WarmupDecayLR(total_num_steps=100, warmup_min_lr=0.0, warmup_max_lr=1.0, warmup_num_steps=10, warmup_type="linear")
    lrs = []
    for _ in range(101):
        s.step()
        lrs.append(s.get_lr()[0])
    assert lrs[0] == 0.0 and abs(lrs[5] - 0.5) < 1e-9 and lrs[10] == 1.0 and abs(lrs[55] - 0.5) < 1e-9 and lrs[100] == 0.0
