"""Checkpoint I/O compatibility with the reference's tooling (offline, never on the hot path).

Replaces / interoperates with:
  * params_bf16_to_f32.py:5-28        merge a DeepSpeed checkpoint directory (``mp_rank_00_model_states.pt`` with the
                                      model under ``"module"`` + one ``layer_{l}_expert_{e}_mp_rank_00_model_states.pt``
                                      per MoE expert) into one fp32 state dict
  * merge_lora_weights_and_save_hf_model(_moe).py:170-188   load that state dict into the peft-wrapped model,
                                      ``merge_and_unload()``, ``save_pretrained``
  * medplib_moe_llama.py:617-635      expert keys ``…mlp.deepspeed_moe.experts.deepspeed_experts.{e}.…``
Our module tree already uses the reference's parameter names, so loading is a key normalisation (``module.`` /
``base_model.model.`` prefixes, peft's ``base_layer``) plus, for LoRA checkpoints, either attaching the adapters or folding
them into the base weights. Plain torch on CPU tensors: this is file plumbing, not compute.
"""
import os
import re

import torch

_EXPERT_RE = re.compile(r"(.*\.)?model\.layers\.(\d+)\.mlp\.deepspeed_moe\.experts\.deepspeed_experts\.(\d+)\.")


def merge_deepspeed_states(directory, dtype=torch.float32, device="cpu"):
    """params_bf16_to_f32.py::load_model_parameters: every ``*model_states.pt`` in `directory`; files with "expert" in
    their name hold the expert tensors directly, the others hold the model under "module". Duplicate keys are an error."""
    combined = {}
    for filename in sorted(os.listdir(directory)):
        if not filename.endswith("model_states.pt"):
            continue
        data = torch.load(os.path.join(directory, filename), map_location=device, weights_only=False)
        sd = data if "expert" in filename else data["module"]
        for k, v in sd.items():
            if k in combined:
                raise ValueError(f"Duplicate key found in state dicts: {k}")
            combined[k] = v.to(dtype) if dtype is not None and torch.is_tensor(v) and v.is_floating_point() else v
    return combined


def peft_state_dict(model):
    """state_dict() of `model` under the key names a peft-wrapped model has (what the reference's
    merge_lora_weights_and_save_hf_model*.py load with strict=False into ``get_peft_model(...)``): a model wrapped by
    medplib_b200.compat.peft_shim (or real peft) already has them; a bare model with attach_lora adapters gets the
    ``base_model.model.`` prefix and ``<linear>.base_layer.weight`` for every adapted Linear."""
    sd = model.state_dict()
    if any(k.startswith("base_model.model.") for k in sd):
        return sd
    adapted = {n for n, m in model.named_modules() if hasattr(m, "lora_A")}
    if not adapted:
        return sd
    out = {}
    for k, v in sd.items():
        mod, _, leaf = k.rpartition(".")
        if mod in adapted and leaf in ("weight", "bias"):
            k = f"{mod}.base_layer.{leaf}"
        out["base_model.model." + k] = v
    return out


def save_deepspeed_layout(model, directory, tag_prefix=None):
    """Write `model` the way DeepSpeed's MoE engine checkpoints it, so the reference's params_bf16_to_f32.py / merge
    scripts / ``engine.load_checkpoint`` can read a model trained here:
      ``mp_rank_00_model_states.pt``  {"module": every non-expert tensor}
      ``layer_{i}_expert_{e}_mp_rank_00_model_states.pt``  the tensors of expert e of the i-th MoE layer -- i counts MoE
      layers (DeepSpeed's moe_layer_id), NOT transformer layers, which matters for --moe_mode second_half / sparse.
    Keys follow peft's naming when adapters are attached (see peft_state_dict); tag_prefix overrides the prefix."""
    os.makedirs(directory, exist_ok=True)
    sd = peft_state_dict(model)
    if tag_prefix:
        sd = {tag_prefix + k: v for k, v in sd.items()}
    rest, experts = {}, {}
    for k, v in sd.items():
        m = _EXPERT_RE.match(k)
        if m:
            experts.setdefault((int(m.group(2)), int(m.group(3))), {})[k] = v.detach().cpu()
        else:
            rest[k] = v.detach().cpu()
    torch.save({"module": rest}, os.path.join(directory, "mp_rank_00_model_states.pt"))
    moe_layer_id = {layer: i for i, layer in enumerate(sorted({l for l, _ in experts}))}
    for (layer, e), esd in experts.items():
        torch.save(esd, os.path.join(directory, f"layer_{moe_layer_id[layer]}_expert_{e}_mp_rank_00_model_states.pt"))
    return len(experts)


def normalize_keys(sd):
    """Strip the wrappers' prefixes (DeepSpeed ``module.``, peft ``base_model.model.``) and peft's ``base_layer``."""
    out = {}
    for k, v in sd.items():
        while True:
            for pre in ("module.", "base_model.model."):
                if k.startswith(pre):
                    k = k[len(pre):]
                    break
            else:
                break
        k = k.replace(".base_layer.", ".")
        out[k] = v
    return out


def lora_keys(sd):
    return sorted(k for k in sd if ".lora_A." in k or ".lora_B." in k)


def fold_lora(sd, scaling):
    """peft merge_and_unload on a state dict: W += scaling * B @ A for every adapted Linear; adapter keys removed."""
    sd = dict(sd)
    for ka in [k for k in sd if k.endswith(".lora_A.default.weight")]:
        base = ka[: -len(".lora_A.default.weight")]
        kb = base + ".lora_B.default.weight"
        w = sd[base + ".weight"]
        sd[base + ".weight"] = (w.float() + scaling * sd[kb].float() @ sd[ka].float()).to(w.dtype)
        del sd[ka], sd[kb]
    return sd


def load_into(model, sd, lora="auto", scaling=None, strict=False):
    """Load a reference-layout state dict (merged DeepSpeed checkpoint, HF ``pytorch_model.bin``, peft-wrapped or not)
    into a MedPLIBForCausalLM. LoRA tensors: ``lora="fold"`` merges them into the base weights (needs `scaling` =
    lora_alpha / r), ``"keep"`` requires adapters to be attached already (medplib_b200.train.attach_lora), ``"auto"``
    keeps them when the model has adapters and folds otherwise. Returns (missing, unexpected) like load_state_dict."""
    sd = normalize_keys(sd)
    has_adapters = any(hasattr(m, "lora_A") for m in model.modules())
    if lora_keys(sd):
        if lora == "fold" or (lora == "auto" and not has_adapters):
            if scaling is None:
                raise ValueError("folding LoRA tensors needs scaling = lora_alpha / r")
            sd = fold_lora(sd, scaling)
        elif not has_adapters:
            raise ValueError("the checkpoint has LoRA tensors but the model has no adapters (attach_lora first)")
    own = model.state_dict()
    # a model whose adapted Linears are peft-style wrappers keeps the frozen matrix under `<name>.base_layer.weight`
    alias = {k.replace(".base_layer.", "."): k for k in own if ".base_layer." in k}
    sd = {alias.get(k, k): v for k, v in sd.items()}
    cast = {}
    for k, v in sd.items():
        if k in own and torch.is_tensor(v) and v.is_floating_point():
            v = v.to(own[k].dtype)
        cast[k] = v
    res = model.load_state_dict(cast, strict=strict)
    if hasattr(model, "refresh_engines"):
        model.refresh_engines()
    return list(res.missing_keys), list(res.unexpected_keys)
