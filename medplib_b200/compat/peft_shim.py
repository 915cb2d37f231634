"""The part of ``peft`` 0.10 that ``train_ds_medplib.py:15,294-302`` and ``merge_lora_weights_and_save_hf_model*.py``
touch — ``LoraConfig``, ``get_peft_model``, a ``PeftModel`` wrapper with peft's module nesting and parameter names —
on top of ``medplib_b200.train.attach_lora``.

Layout reproduced (peft/tuners/lora/layer.py::Linear, peft/peft_model.py::PeftModelForCausalLM):

    PeftModelForCausalLM.base_model (LoraModel) .model (MedPLIBForCausalLM)
        …self_attn.q_proj                      LoraLinear (was nn.Linear)
        …self_attn.q_proj.base_layer           the original nn.Linear  -> `…q_proj.base_layer.weight`
        …self_attn.q_proj.lora_A.default       nn.Linear(in, r, bias=False)   kaiming-uniform(a=sqrt 5)
        …self_attn.q_proj.lora_B.default       nn.Linear(r, out, bias=False)  zeros
        …self_attn.q_proj.lora_dropout.default nn.Dropout(p) (nn.Identity when p == 0)

so ``state_dict()`` keys read ``base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight`` exactly as a
checkpoint written under real peft, and ``--sft_modules`` / LoRA-target substring matching keeps working. The adapter
arithmetic itself (y = W x + (alpha / r) B A dropout(x)) runs in medplib_b200's train step, not here.
"""
import torch
import torch.nn as nn

from .. import _lib
from .. import train as _train


class LoraConfig:
    def __init__(self, r=8, lora_alpha=8, target_modules=None, lora_dropout=0.0, bias="none", task_type=None,
                 **unused):
        if bias != "none":
            raise _lib.MplError('LoraConfig(bias=...) other than "none" is not built (the reference uses "none")')
        self.r, self.lora_alpha, self.lora_dropout = int(r), lora_alpha, float(lora_dropout)
        self.target_modules = list(target_modules) if not isinstance(target_modules, str) else target_modules.split(",")
        self.bias, self.task_type = bias, task_type
        self.peft_type = "LORA"


class LoraLinear(nn.Module):
    """peft's lora.Linear as a parameter holder: medplib_b200's kernels read ``base_layer.weight`` and the adapter
    matrices in place (``medplib_b200/train.py:_Lora``); calling it is an error like every module of this build."""

    def __init__(self, base_layer, r, lora_alpha, lora_dropout):
        super().__init__()
        w = base_layer.weight
        self.base_layer = base_layer
        self.in_features, self.out_features = base_layer.in_features, base_layer.out_features
        A = nn.Linear(self.in_features, r, bias=False, device=w.device, dtype=w.dtype)
        B = nn.Linear(r, self.out_features, bias=False, device=w.device, dtype=w.dtype)
        nn.init.kaiming_uniform_(A.weight, a=5 ** 0.5)
        nn.init.zeros_(B.weight)
        self.lora_A = nn.ModuleDict({"default": A})
        self.lora_B = nn.ModuleDict({"default": B})
        self.lora_dropout = nn.ModuleDict({"default": nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()})
        self.r, self.lora_alpha = {"default": r}, {"default": lora_alpha}
        self.scaling = {"default": lora_alpha / r}
        self.lora_dropout_p = float(lora_dropout)
        self.active_adapter = "default"
        self.merged = False

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, *a, **k):
        raise _lib.MplError("medplib_b200 modules only hold parameters; the adapter runs inside the fused train step")


def _wrap_targets(model, cfg):
    names = _train.find_linear_layers(model, cfg.target_modules)
    mods = dict(model.named_modules())
    for name in names:
        parent_name, _, child = name.rpartition(".")
        parent = mods[parent_name] if parent_name else model
        lin = getattr(parent, child)
        if isinstance(lin, LoraLinear):
            continue
        wrapped = LoraLinear(lin, cfg.r, cfg.lora_alpha, cfg.lora_dropout)
        wrapped.lora_name = name
        setattr(parent, child, wrapped)
    return names


class LoraModel(nn.Module):
    def __init__(self, model, cfg):
        super().__init__()
        self.model = model
        self.peft_config = {"default": cfg}
        for p in model.parameters():  # peft: mark_only_lora_as_trainable
            p.requires_grad = False
        self.targets = _wrap_targets(model, cfg)
        for n, p in model.named_parameters():
            if "lora_" in n:
                p.requires_grad = True
        if hasattr(model, "refresh_engines"):
            model.refresh_engines()

    def forward(self, *a, **k):
        return self.model(*a, **k)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.model, name)

    def merge_and_unload(self):
        """Fold every adapter into its base weight and restore the plain nn.Linear modules (peft semantics)."""
        model = self.model
        mods = dict(model.named_modules())
        with torch.no_grad():
            for name, mod in list(mods.items()):
                if not isinstance(mod, LoraLinear):
                    continue
                A, B = mod.lora_A["default"].weight, mod.lora_B["default"].weight
                w = mod.base_layer.weight
                w.add_((mod.scaling["default"] * (B.float() @ A.float())).to(w.dtype))
                parent_name, _, child = name.rpartition(".")
                setattr(mods[parent_name] if parent_name else model, child, mod.base_layer)
        if hasattr(model, "refresh_engines"):
            model.refresh_engines()
        return model


class PeftModel(nn.Module):
    def __init__(self, model, peft_config, adapter_name="default"):
        super().__init__()
        self.base_model = LoraModel(model, peft_config)
        self.peft_config = {adapter_name: peft_config}
        self.active_adapter = adapter_name
        self.config = getattr(model, "config", None)

    def forward(self, *a, **k):
        return self.base_model(*a, **k)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.base_model, name)

    def get_base_model(self):
        return self.base_model.model

    def get_nb_trainable_parameters(self):
        tr = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return tr, sum(p.numel() for p in self.parameters())

    def print_trainable_parameters(self):
        tr, tot = self.get_nb_trainable_parameters()
        print(f"trainable params: {tr:,d} || all params: {tot:,d} || trainable%: {100 * tr / max(tot, 1):.4f}")

    def merge_and_unload(self):
        return self.base_model.merge_and_unload()

    def save_pretrained(self, path, **kw):
        import os
        os.makedirs(path, exist_ok=True)
        sd = {k.replace(".default", ""): v for k, v in self.state_dict().items() if "lora_" in k}
        torch.save(sd, os.path.join(path, "adapter_model.bin"))


class PeftModelForCausalLM(PeftModel):
    def generate(self, *a, **k):
        return self.base_model.model.generate(*a, **k)


def get_peft_model(model, peft_config, adapter_name="default"):
    return PeftModelForCausalLM(model, peft_config, adapter_name)
