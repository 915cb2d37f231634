"""CPU: the reference's OWN training driver (`/root/reference/train_ds_medplib.py`, unmodified, imported from where it
lies) runs its whole construction sequence against medplib_b200's classes — argument parsing, tokenizer surgery,
``from_pretrained(**vars(args))``, ``initialize_vision_modules`` / ``initialize_bird_modules`` /
``initialize_lisa_modules``, LoRA target discovery + ``get_peft_model``, ``initialize_moe_modules``,
``resize_token_embeddings``, the ``--sft_modules`` loop, the DeepSpeed config + ``deepspeed.initialize``, the epoch loop
and ``save_checkpoint`` — with only what a GPU-less, network-less box cannot provide replaced:

  * ``model.MedPLIB`` / ``model.LISA``   -> medplib_b200.model (the two import lines of INTEGRATION.md §2)
  * ``peft`` / ``deepspeed``            -> medplib_b200.compat (neither package is installable here)
  * tokenizer / dataset / TensorBoard   -> small stand-ins (no tokenizer files, no images on disk)
  * ``train()`` / ``validate()``         -> recorded (the forward needs a B200: tests/test_train_gpu.py)
  * ``Module.to(device=<int>)`` / ``torch.cuda.device_count()``  -> CPU / 1 (the driver moves the tower to ``local_rank``)

Skipped when /root/reference is absent (the GPU box)."""
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_ds_medplib.py")),
                                reason="needs the reference tree")
CLIP_CFG = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, image_size=56,
                patch_size=14, layer_norm_eps=1e-5)


class StubTokenizer:
    """What the driver needs of the LLaMA tokenizer: ids for added tokens, len(), special ids."""

    def __init__(self, base_vocab=300):
        self.vocab = {f"tok{i}": i for i in range(base_vocab)}
        self.unk_token, self.pad_token = "<unk>", None
        self.eos_token_id, self.bos_token_id, self.unk_token_id = 2, 1, 0
        self.model_max_length = 512

    @property
    def pad_token_id(self):
        return self.unk_token_id

    def add_tokens(self, toks, special_tokens=False):
        toks = [toks] if isinstance(toks, str) else toks
        n = 0
        for t in toks:
            if t not in self.vocab:
                self.vocab[t] = len(self.vocab)
                n += 1
        return n

    def __call__(self, text, add_special_tokens=True):
        return types.SimpleNamespace(input_ids=[self.vocab[text]])

    def __len__(self):
        return len(self.vocab)


class TinyDataset(torch.utils.data.Dataset):
    def __init__(self, *a, **k):
        pass

    def __len__(self):
        return 8

    def __getitem__(self, i):
        return {"i": i}


def _make_checkpoint(tmp, moe):
    from medplib_b200.model import LISAForCausalLM, MedPLIBForCausalLM, MedPLIBMoELlamaConfig
    from medplib_b200.model.config import LlavaConfig
    torch.manual_seed(0)
    kw = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2,
              vocab_size=300, rms_norm_eps=1e-5, max_position_embeddings=512, mm_vision_select_layer=-2,
              mm_projector_type="mlp2x_gelu", max_sample_point=512)
    if moe:
        cfg = MedPLIBMoELlamaConfig(**kw)
        cls = MedPLIBForCausalLM
    else:
        cfg = LlavaConfig(**kw)
        cls = LISAForCausalLM
    cfg.clip_config = CLIP_CFG
    cfg.sam_config = dict(image_size=256, embed_dim=32, depth=1, num_heads=2)
    m = cls(cfg, seg_token_idx=299, train_mask_decoder=True, out_dim=256)
    path = os.path.join(tmp, "ckpt_moe" if moe else "ckpt_dense")
    m.save_pretrained(path)
    # a SAM-Med2D checkpoint in its own format ({"model": state_dict}, build_sam.py:123-148) with recognisable values
    g = torch.Generator().manual_seed(7)
    sam_sd = {k: torch.randn(v.shape, generator=g) * 0.01 for k, v in m.model.visual_model.state_dict().items()}
    torch.save({"model": sam_sd}, os.path.join(tmp, "sam_med2d.pth"))
    return path


@pytest.fixture()
def driver(monkeypatch, tmp_path):
    """Import the unmodified driver with the substitutions listed in the module docstring."""
    import medplib_b200.compat as compat
    import medplib_b200.model as ours
    for name in [n for n in sys.modules if n == "model" or n.startswith("model.") or n == "datasets"
                 or n.startswith("datasets.") or n == "utils" or n.startswith("utils.") or n == "train_ds_medplib"]:
        monkeypatch.delitem(sys.modules, name, raising=False)
    monkeypatch.syspath_prepend(REF)
    installed = compat.install(force=True)
    assert set(installed) == {"peft", "deepspeed"}
    for mod, cls in (("model.MedPLIB", "MedPLIBForCausalLM"), ("model.LISA", "LISAForCausalLM")):
        stub = types.ModuleType(mod)
        setattr(stub, cls, getattr(ours, cls))
        monkeypatch.setitem(sys.modules, mod, stub)
    import transformers
    monkeypatch.setattr(transformers.AutoTokenizer, "from_pretrained", classmethod(lambda cls, *a, **k: StubTokenizer()))
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 1)
    orig_to = nn.Module.to

    def to_cpu_when_no_gpu(self, *a, **k):  # vision_tower.to(dtype=..., device=args.local_rank)
        if isinstance(k.get("device"), int) and not torch.cuda.is_available():
            k["device"] = "cpu"
        return orig_to(self, *a, **k)

    monkeypatch.setattr(nn.Module, "to", to_cpu_when_no_gpu)
    import torch.utils.tensorboard as tb
    monkeypatch.setattr(tb, "SummaryWriter", lambda *a, **k: types.SimpleNamespace(add_scalar=lambda *a, **k: None))
    import train_ds_medplib as drv
    monkeypatch.setattr(drv, "LazySupervisedDataset", TinyDataset)
    monkeypatch.setattr(drv, "ICLLazySupervisedDataset", TinyDataset)
    calls = {"train": [], "validate": []}

    def fake_train(train_loader, model, epoch, scheduler, writer, train_iter, args):
        calls["train"].append(dict(engine=model, loader=train_loader, scheduler=scheduler, epoch=epoch, args=args))
        return train_iter

    monkeypatch.setattr(drv, "train", fake_train)
    monkeypatch.setattr(drv, "validate", lambda *a, **k: (0.0, 0.0))
    monkeypatch.setattr(torch.distributed, "barrier", lambda *a, **k: None)
    yield drv, calls, str(tmp_path)
    for name in ("peft", "deepspeed", "deepspeed.moe", "deepspeed.moe.layer", "deepspeed.moe.utils"):
        sys.modules.pop(name, None)


def _argv(ckpt, tmp, extra):
    return ["--version", ckpt, "--vision_tower", "random-clip", "--precision", "bf16", "--log_base_dir", tmp,
            "--exp_name", "exp", "--epochs", "1", "--batch_size", "2", "--grad_accumulation_steps", "2", "--no_eval",
            "--lora_r", "8", "--lora_alpha", "16", "--lora_dropout", "0.05", "--train_mask_decoder",
            "--data_path", "unused.json", "--val_data_path", "unused.json",
            "--vision_pretrained", os.path.join(tmp, "sam_med2d.pth")] + extra


def test_stage4_moe_recipe_constructs_through_the_unmodified_driver(driver):
    """scripts/train_stage4.sh flags: --moe_enable, LoRA on q,v,gate,up,down, sft on wg / lm_head / embed_tokens /
    mask_decoder / text_hidden_fcs / region_fea_adapter."""
    drv, calls, tmp = driver
    ckpt = _make_checkpoint(tmp, moe=True)
    drv.main(_argv(ckpt, tmp, ["--moe_enable", "True", "--moe_mode", "dense", "--num_experts", "2", "--top_k_experts", "1",
                               "--capacity_factor", "1.5", "--router_aux_loss_coef", "0.0", "--region_fea_adapter",
                               "--lora_target_modules", "q_proj,v_proj,gate_proj,up_proj,down_proj", "--sft_modules",
                               "wg,lm_head,embed_tokens,mask_decoder,text_hidden_fcs,region_fea_adapter"]))
    assert len(calls["train"]) == 1
    c = calls["train"][0]
    eng, args = c["engine"], c["args"]
    names = dict(eng.module.named_parameters())
    trainable = {n for n, p in names.items() if p.requires_grad}
    # peft's nesting and names (what merge_lora_weights_and_save_hf_model_moe.py expects in a checkpoint)
    assert "base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight" in trainable
    assert "base_model.model.model.layers.0.self_attn.q_proj.base_layer.weight" in names
    assert "base_model.model.model.layers.0.self_attn.q_proj.base_layer.weight" not in trainable
    e0 = "base_model.model.model.layers.1.mlp.deepspeed_moe.experts.deepspeed_experts.1."
    assert e0 + "up_proj.lora_B.default.weight" in trainable and e0 + "down_proj.base_layer.weight" in names
    assert "base_model.model.model.layers.0.mlp.deepspeed_moe.gate.wg.weight" in trainable
    assert "base_model.model.lm_head.weight" in trainable and "base_model.model.model.embed_tokens.weight" in trainable
    assert any("mask_decoder" in n for n in trainable) and any("text_hidden_fcs" in n for n in trainable)
    assert any("region_fea_adapter" in n for n in trainable)
    assert not any("vision_tower" in n or "image_encoder" in n or "mm_projector" in n for n in trainable)
    # initialize_bird_modules rebuilt SAM-Med2D from --vision_pretrained (build_sam.py:123-148 format)
    sam_sd = torch.load(os.path.join(tmp, "sam_med2d.pth"))["model"]
    k = "image_encoder.blocks.0.attn.qkv.weight"
    got = names["base_model.model.model.visual_model." + k]
    assert torch.allclose(got.float(), sam_sd[k].to(got.dtype).float())
    # resize_token_embeddings(len(tokenizer)): 300 base + 265 ADD_OTHERS + 2 im_start/end (utils/utils.py:16)
    assert names["base_model.model.lm_head.weight"].shape[0] == args.seg_token_idx + 1 or \
        names["base_model.model.lm_head.weight"].shape[0] > 300
    # the DeepSpeed config reached the engine: micro-batch 2, accumulation 2, AdamW(0.9, 0.95), clip 1, WarmupDecayLR
    assert eng.gradient_accumulation_steps() == 2 and eng._micro_bs == 2
    assert eng._opt_kw["betas"] == (0.9, 0.95) and eng._opt_kw["max_grad_norm"] == 1.0
    assert type(c["scheduler"]).__name__ == "WarmupDecayLR"
    assert c["loader"].batch_size == 2 and c["loader"].collate_fn.func.__name__ == "DataCollatorForSupervisedDataset"
    # end of epoch: engine.save_checkpoint(<log_dir>/last_ckpt_model) in the layout params_bf16_to_f32.py merges
    out = os.path.join(tmp, "exp", "last_ckpt_model")
    tag = open(os.path.join(out, "latest")).read().strip()
    files = sorted(os.listdir(os.path.join(out, tag)))
    assert "mp_rank_00_model_states.pt" in files
    assert sum(f.startswith("layer_") and "_expert_" in f for f in files) == 2 * 2


def test_dense_recipe_goes_through_initialize_lisa_modules(driver):
    """Stage 1-3 recipes (no --moe_enable): LISAForCausalLM + initialize_lisa_modules (train_ds_medplib.py:247)."""
    drv, calls, tmp = driver
    ckpt = _make_checkpoint(tmp, moe=False)
    drv.main(_argv(ckpt, tmp, ["--lora_target_modules", "q_proj,v_proj", "--sft_modules",
                               "lm_head,embed_tokens,mask_decoder,text_hidden_fcs"]))
    assert len(calls["train"]) == 1
    eng = calls["train"][0]["engine"]
    trainable = {n for n, p in eng.module.named_parameters() if p.requires_grad}
    assert "base_model.model.model.layers.1.self_attn.v_proj.lora_B.default.weight" in trainable
    assert not any("gate_proj" in n for n in trainable)
    assert any("text_hidden_fcs" in n for n in trainable) and any("mask_decoder" in n for n in trainable)
