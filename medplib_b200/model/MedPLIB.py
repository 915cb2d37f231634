"""MedPLIBForCausalLM — drop-in for the reference's model/MedPLIB.py::MedPLIBForCausalLM (:187-701) whose arithmetic
runs in libmedplib_b200.so (hand-written sm_100a kernels) instead of transformers / DeepSpeed / torch eager ops.

Kept from the reference: class name, constructor kwargs (MedPLIB.py:195-247), module tree and
parameter names, ``forward(**input_dict)`` dispatch (:359-362), ``model_forward`` signature and returned dict keys
(:364-572), ``evaluate`` (:574-680), ``generate`` keyword surface (model/eval/vqa_infer.py:430-442),
``get_visual_embs`` / ``postprocess_masks`` / ``build_seg_token_mask`` / ``expand_embedding`` / ``encode_images`` /
``get_model().get_vision_tower()`` / ``initialize_*`` methods.
Deliberate fixes of reference quirks, each listed in SURVEY.md App. B: B-1 (seg mask truncated to the hidden length),
B-10 (no swallowed exception around the SAM encoder), B-12 (``inference=True`` usable with the MoE wrapper), B-13 (each
sub-module built once). Work the reference does and discards is skipped without changing results: the 24th CLIP layer,
``text_hidden_fcs`` on non-[SEG] rows, the three unused hypernetwork MLPs, fp32 logits for non-final positions during
generation.
There is no eager fallback: without the built library or on a CPU tensor every entry point raises MplError.
"""
import os
import types
import weakref
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
from transformers import PreTrainedModel
from transformers.utils import ModelOutput

from .. import _lib, engine, ops
from . import modules as M
from .config import MedPLIBMoELlamaConfig, llama_dims

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
REGION_TOKEN_INDEX = -300
bf16 = torch.bfloat16

CLIP_L_336 = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                  image_size=336, patch_size=14, layer_norm_eps=1e-5)


@dataclass
class MoECausalLMOutputWithPast(ModelOutput):
    loss: Optional[torch.FloatTensor] = None
    moe_loss: Optional[torch.FloatTensor] = None
    logits: torch.FloatTensor = None
    past_key_values: Optional[object] = None
    hidden_states: Optional[Tuple[torch.FloatTensor]] = None
    attentions: Optional[Tuple[torch.FloatTensor]] = None
    moe_loss_list: Optional[Tuple[torch.FloatTensor]] = None


@dataclass
class GenerateOutput(ModelOutput):
    sequences: torch.LongTensor = None
    hidden_states: Optional[Tuple] = None
    last_hidden_state: Optional[torch.Tensor] = None  # [B, P_spliced + new - 1, D]: concat of per-step last-layer states
    scores: Optional[Tuple] = None  # per-step fp32 logits [B, V] when output_scores=True


class CLIPVisionTower(nn.Module):
    """model/medplib/model/multimodal_encoder/clip_encoder.py:6-87 (frozen tower, hidden_states[select_layer], no CLS)."""

    def __init__(self, vision_tower, args, delay_load=False):
        super().__init__()
        self.is_loaded = False
        self.vision_tower_name = vision_tower
        self.select_layer = getattr(args, "mm_vision_select_layer", -2)
        self.select_feature = getattr(args, "mm_vision_select_feature", "patch")
        self._engine = None
        self._clip_config = getattr(args, "clip_config", None)  # dict: random-init tower of a given size (tests)
        if not delay_load:
            self.load_model()

    def load_model(self):
        from transformers import CLIPImageProcessor, CLIPVisionConfig
        name = self.vision_tower_name
        if name and os.path.isdir(str(name)):
            cfg = CLIPVisionConfig.from_pretrained(name)
            self.image_processor = CLIPImageProcessor.from_pretrained(name)
        else:  # no network in this build: random-init tower of the reference architecture
            cfg = CLIPVisionConfig(**(self._clip_config or CLIP_L_336))
            self.image_processor = CLIPImageProcessor(size={"shortest_edge": cfg.image_size},
                                                      crop_size={"height": cfg.image_size, "width": cfg.image_size})
        self.vision_tower = M.CLIPVisionModelHolder(cfg)
        if name and os.path.isdir(str(name)):
            _load_checkpoint_into(self.vision_tower, name)
        self.vision_tower.requires_grad_(False)
        self.is_loaded = True
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def forward(self, images):
        if self.select_feature != "patch":
            raise _lib.MplError("only mm_vision_select_feature='patch' is built (the reference's configuration)")
        if self._engine is None:
            c = self.config
            cfg = dict(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                       num_layers=c.num_hidden_layers, num_heads=c.num_attention_heads, image_size=c.image_size,
                       patch_size=c.patch_size, layer_norm_eps=c.layer_norm_eps)
            sd = {k: v for k, v in self.vision_tower.state_dict(keep_vars=True).items()}
            for v in sd.values():
                if v.is_floating_point() and v.dtype != bf16:
                    raise _lib.MplError("the vision tower must be bf16 on a CUDA device (model.to(torch.bfloat16).cuda())")
            self._engine = engine.ClipEngine(sd, cfg, "", self.select_layer)
        if isinstance(images, list):
            return [self._engine.forward(im.unsqueeze(0)) for im in images]
        return self._engine.forward(images)

    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return self.vision_tower.vision_model.pre_layrnorm.weight.dtype

    @property
    def device(self):
        return self.vision_tower.vision_model.pre_layrnorm.weight.device

    @property
    def config(self):
        return self.vision_tower.config

    @property
    def hidden_size(self):
        return self.config.hidden_size

    @property
    def num_patches(self):
        return (self.config.image_size // self.config.patch_size) ** 2


def _load_checkpoint_into(module, path):
    """Load *.safetensors / *.bin found in `path` into `module` (non-strict)."""
    import glob
    sd = {}
    for f in sorted(glob.glob(os.path.join(path, "*.safetensors"))):
        from safetensors.torch import load_file
        sd.update(load_file(f))
    if not sd:
        for f in sorted(glob.glob(os.path.join(path, "*.bin"))):
            sd.update(torch.load(f, map_location="cpu"))
    return module.load_state_dict(sd, strict=False)


class MedPLIBModel(nn.Module):
    """``self.model`` of the reference: LlamaModel + LlavaMetaModel (medplib_arch.py:111-187) + MedPLIBMetaModel
    (MedPLIB.py:127-164). Same attribute / parameter names."""

    def __init__(self, config, **kwargs):
        super().__init__()
        self.config = config
        D = config.hidden_size
        self.padding_idx = getattr(config, "pad_token_id", None)
        self.vocab_size = config.vocab_size
        self.embed_tokens = nn.Embedding(config.vocab_size, D)
        self.layers = nn.ModuleList([M.LlamaDecoderLayer(config) for _ in range(config.num_hidden_layers)])
        self.norm = M.LlamaRMSNorm(D, eps=config.rms_norm_eps)
        self.gradient_checkpointing = False
        # LlavaMetaModel
        self.vision_tower = CLIPVisionTower(getattr(config, "mm_vision_tower", getattr(config, "vision_tower", None)),
                                            config)
        Dv = getattr(config, "mm_hidden_size", self.vision_tower.hidden_size)
        config.mm_hidden_size = Dv
        ptype = getattr(config, "mm_projector_type", "mlp2x_gelu")
        if ptype == "mlp2x_gelu":
            self.mm_projector = nn.Sequential(nn.Linear(Dv, D), nn.GELU(), nn.Linear(D, D))
        elif ptype == "linear":
            self.mm_projector = nn.Linear(Dv, D)
        else:
            raise _lib.MplError(f"mm_projector_type {ptype!r} is not built (reference checkpoints use mlp2x_gelu)")
        if getattr(config, "mm_token_compress", False):
            self.mm_token_compressor = M.TokenCompressor(D, getattr(config, "mm_compressed_token_count", 256))
        if getattr(config, "icl_mask_encoder", False):
            self.mask_encoder = M.MaskTokenEncoder(D, getattr(config, "mask_encoder_token_count", 64))
        self.region_fea_adapter = nn.Linear(Dv, D)
        if getattr(config, "region_geo_sampler", False):
            # the instantiation the reference keeps commented out (medplib_arch.py:134-142), with its arguments
            from .geo_sampler import GeoRegionSampler
            self.region_geo_sampler = GeoRegionSampler(input_dim=Dv, output_dim=D,
                                                       num_init_point=getattr(config, "max_sample_point", 512) or 512,
                                                       num_sub_point=[128, 32], num_neighbor=[24, 24],
                                                       pooler_mode=getattr(config, "sampler_pooler_mode", None) or "max")
        self.max_sample_point = getattr(config, "max_sample_point", 512)
        # MedPLIBMetaModel
        self.vision_pretrained = kwargs.get("vision_pretrained", None)
        self.initialize_bird_modules(config)
        config.use_cache = False
        config.vision_tower = getattr(config, "mm_vision_tower", None)
        config.mm_vision_select_feature = "patch"
        config.image_aspect_ratio = "square"
        config.image_grid_pinpoints = None
        config.tune_mm_mlp_adapter = False
        config.freeze_mm_mlp_adapter = True
        config.mm_use_im_patch_token = False

    def get_vision_tower(self):
        vt = getattr(self, "vision_tower", None)
        return vt[0] if type(vt) is list else vt

    def initialize_vision_modules(self, model_args, fsdp=None):
        """medplib_arch.py:150-187. Re-points the tower at model_args.vision_tower and (optionally) loads the projector."""
        vt = model_args.vision_tower
        self.config.mm_vision_tower = vt
        cur = self.get_vision_tower()
        if cur is None or cur.vision_tower_name != vt or not cur.is_loaded:
            dev, dt = (cur.device, cur.dtype) if cur is not None and cur.is_loaded else (None, None)
            tower = CLIPVisionTower(vt, model_args)
            if dev is not None:
                tower.to(device=dev, dtype=dt)
            self.vision_tower = tower
        self.config.use_mm_proj = True
        self.config.mm_hidden_size = self.get_vision_tower().hidden_size
        self.config.mm_vision_select_layer = getattr(model_args, "mm_vision_select_layer", -2)
        self.config.mm_vision_select_feature = getattr(model_args, "mm_vision_select_feature", "patch")
        path = getattr(model_args, "pretrain_mm_mlp_adapter", None)
        if path is not None:
            w = torch.load(path, map_location="cpu")
            self.mm_projector.load_state_dict({k.split("mm_projector.")[1]: v for k, v in w.items()
                                               if "mm_projector" in k})

    def initialize_bird_modules(self, config):
        """MedPLIB.py:141-164: (re)build SAM-Med2D from ``vision_pretrained`` (frozen; the mask decoder trainable when
        train_mask_decoder) and a fresh text_hidden_fcs. Like the reference, an explicit call from the train driver
        (train_ds_medplib.py:245) REPLACES both — weights that from_pretrained put there are dropped."""
        sam_cfg = getattr(config, "sam_config", None) or {}
        ref = next(self.parameters(), None)
        self.visual_model = M.Sam(**sam_cfg)
        if self.vision_pretrained is not None:
            # build_sam.py:123-148: {"model": state_dict} checkpoints (SAM-Med2D's own) load non-strictly, bare state
            # dicts strictly (image_size 256 with adapters)
            with open(self.vision_pretrained, "rb") as f:
                state = torch.load(f, map_location="cpu")
            if "model" in state:
                self.visual_model.load_state_dict(state["model"], strict=False)
            else:
                self.visual_model.load_state_dict(state)
        in_dim, out_dim = config.hidden_size, getattr(config, "out_dim", 256)
        self.text_hidden_fcs = nn.ModuleList([nn.Sequential(
            nn.Linear(in_dim, in_dim), nn.ReLU(inplace=True), nn.Linear(in_dim, out_dim), nn.Dropout(0.0))])
        if ref is not None and (ref.is_cuda or ref.dtype != torch.float32) and hasattr(self, "embed_tokens"):
            # rebuilt on an already placed model (the driver calls this after from_pretrained): follow it
            self.visual_model.to(device=ref.device, dtype=ref.dtype)
            self.text_hidden_fcs.to(device=ref.device, dtype=ref.dtype)
        for p in self.visual_model.parameters():
            p.requires_grad = False
        if getattr(config, "train_mask_decoder", True):
            self.visual_model.mask_decoder.train()
            for p in self.visual_model.mask_decoder.parameters():
                p.requires_grad = True
        self.text_hidden_fcs.train()
        for p in self.text_hidden_fcs.parameters():
            p.requires_grad = True
        owner = getattr(self, "_owner", None)
        if owner is not None and owner() is not None:
            owner().refresh_engines()

    def initialize_lisa_modules(self, config):
        """model/LISA.py:134-158 — the non-MoE recipes (train_ds_medplib.py:247,
        merge_lora_weights_and_save_hf_model.py:109) call the same construction under this name."""
        return self.initialize_bird_modules(config)

    def forward(self, *a, **k):
        raise _lib.MplError("call MedPLIBForCausalLM (the decoder stack runs as one fused native call)")


def _moe_layers_idx(moe_mode, num_layers):
    if moe_mode == "first_half":
        return list(range(0, num_layers // 2))
    if moe_mode == "second_half":
        return list(range(num_layers // 2, num_layers))
    if moe_mode == "sparse":
        return list(range(num_layers))[::2]
    if moe_mode == "dense":
        return list(range(num_layers))
    raise NotImplementedError(f'Only support ["first_half", "second_half", "sparse", "dense"], but found {moe_mode}')


class MedPLIBForCausalLM(PreTrainedModel):
    config_class = MedPLIBMoELlamaConfig
    base_model_prefix = "model"
    supports_gradient_checkpointing = True
    _no_split_modules = ["LlamaDecoderLayer"]

    def __init__(self, config, **kwargs):
        if not kwargs.get("test_only", False):
            config.mm_use_im_start_end = kwargs.pop("use_mm_start_end", True)
            config.train_mask_decoder = kwargs.get("train_mask_decoder", True)
            config.out_dim = kwargs.get("out_dim", 256)
            config.moe = {k: kwargs.get(k) for k in (
                "num_experts", "top_k_experts", "capacity_factor", "use_residual", "router_aux_loss_coef",
                "eval_capacity_factor", "moe_layers_idx", "min_capacity", "ep_size")}
        config.icl_mask_encoder = kwargs.get("icl_mask_encoder", getattr(config, "icl_mask_encoder", False))
        config.mask_encoder_token_count = kwargs.get("mask_encoder_token_count",
                                                     getattr(config, "mask_encoder_token_count", 64))
        config.mm_token_compress = kwargs.get("mm_token_compress", getattr(config, "mm_token_compress", False))
        config.mm_compressed_token_count = kwargs.get("mm_compressed_token_count",
                                                      getattr(config, "mm_compressed_token_count", 256))
        self.ce_loss_weight = kwargs.pop("ce_loss_weight", None)
        self.dice_loss_weight = kwargs.pop("dice_loss_weight", None)
        self.bce_loss_weight = kwargs.pop("bce_loss_weight", None)
        self.iou_loss_weight = kwargs.pop("iou_loss_weight", None)
        self.focal_loss_weight = kwargs.pop("focal_loss_weight", None)
        self.seg_token_idx = kwargs.pop("seg_token_idx", None)
        super().__init__(config)
        self.model = MedPLIBModel(config, **kwargs)
        object.__setattr__(self.model, "_owner", weakref.ref(self))  # module rebuilds invalidate the native tables
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.router_aux_loss_coef = (getattr(config, "moe", None) or {}).get("router_aux_loss_coef", 0.0) or 0.0
        self._eng = {}
        self.post_init()
        if kwargs.get("test_only", False):
            self._build_moe_layers(config.moe)

    # ------------------------------------------------------------------ HF plumbing
    def _init_weights(self, module):
        """Random init of the tensors a checkpoint did not provide. from_pretrained() calls this AFTER loading: tensors
        that came from the checkpoint carry `_is_hf_initialized` and must be left alone."""
        std = getattr(self.config, "initializer_range", 0.02)

        def fresh(t):
            return t is not None and not getattr(t, "_is_hf_initialized", False)

        if isinstance(module, (nn.Linear, nn.Conv2d, nn.ConvTranspose2d)):
            if fresh(module.weight):
                module.weight.data.normal_(mean=0.0, std=std)
            if fresh(module.bias):
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            if fresh(module.weight):
                module.weight.data.normal_(mean=0.0, std=std)

    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.get_model().get_vision_tower()

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def set_input_embeddings(self, value):
        self.model.embed_tokens = value

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new):
        self.lm_head = new

    def _apply(self, fn, *a, **k):
        self._eng = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._eng = {}
        return super().load_state_dict(*a, **k)

    def refresh_engines(self):
        """Rebuild the native weight tables (call after replacing parameter tensors by hand)."""
        self._eng = {}
        vt = self.get_vision_tower()
        if vt is not None:
            vt._engine = None

    # ------------------------------------------------------------------ MoE construction
    def _build_moe_layers(self, moe):
        n = self.config.num_hidden_layers
        idx = moe.get("moe_layers_idx")
        if idx is None:
            idx = _moe_layers_idx(moe.get("moe_mode", "dense"), n)
        nexp = moe["num_experts"]
        if len(nexp) == 1:
            nexp = list(nexp) * len(idx)
        moe["moe_layers_idx"], moe["num_experts"] = list(idx), list(nexp)
        for e, l in zip(nexp, idx):
            layer = self.model.layers[l]
            if isinstance(layer.mlp, M.MoE):
                continue
            layer.mlp = M.MoE(self.config.hidden_size, expert=layer.mlp, num_experts=e, ep_size=moe.get("ep_size", 1),
                              k=moe["top_k_experts"], capacity_factor=moe["capacity_factor"],
                              eval_capacity_factor=moe["eval_capacity_factor"], min_capacity=moe["min_capacity"],
                              use_residual=moe.get("use_residual", False))
        self._eng = {}

    def initialize_moe_modules(self, model_args):
        """medplib_moe_llama.py:488-649: turn the selected layers' MLPs into MoE layers whose experts start as copies of
        the dense MLP, optionally loading per-expert weights from model_args.expert_pretrained_path."""
        import glob
        paths = [p for p in (getattr(model_args, "expert_pretrained_path", "") or "").split(",") if p]
        expert_sd = {}
        for i, p in enumerate(paths):
            assert os.path.exists(p), f"{p} does not exist"
            tmp = {}
            for f in glob.glob(os.path.join(p, "*.bin")):
                tmp.update(torch.load(f, map_location="cpu"))
            expert_sd[i] = {k: v for k, v in tmp.items() if any(t in k for t in ("gate_proj", "up_proj", "down_proj"))}
            targets = (("text_hidden_fcs", "model.text_hidden_fcs.", self.model.text_hidden_fcs),
                       ("mask_decoder", "model.visual_model.mask_decoder.", self.model.visual_model.mask_decoder)) \
                if i == 0 else (("region_fea_adapter", "model.region_fea_adapter.", self.model.region_fea_adapter),)
            for key, prefix, mod in targets:
                mod.load_state_dict({k.replace(prefix, ""): v for k, v in tmp.items() if key in k}, strict=False)
        moe = self.config.moe
        for k in ("moe_enable", "moe_mode", "moe_layers_idx", "ep_size", "top_k_experts", "capacity_factor",
                  "eval_capacity_factor", "min_capacity", "use_residual", "router_aux_loss_coef"):
            moe[k] = getattr(model_args, k, moe.get(k))
        self.router_aux_loss_coef = moe["router_aux_loss_coef"]
        moe["num_experts"] = list(model_args.num_experts)
        if moe["moe_layers_idx"] is None:
            moe["moe_layers_idx"] = _moe_layers_idx(moe["moe_mode"], self.config.num_hidden_layers)
        self._build_moe_layers(moe)
        for l in moe["moe_layers_idx"]:
            for e_idx, e in enumerate(self.model.layers[l].mlp.deepspeed_moe.experts.deepspeed_experts):
                if e_idx in expert_sd:
                    src = expert_sd[e_idx]
                    e.load_state_dict({f"{n}.weight": src[f"model.layers.{l}.mlp.{n}.weight"]
                                       for n in ("gate_proj", "up_proj", "down_proj")}, strict=False)

    # ------------------------------------------------------------------ engines
    def _params(self):
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        # peft's lora.Linear keeps the frozen matrix as `<name>.base_layer.weight`: the native tables look it up as
        # `<name>.weight` (medplib_b200/compat/peft_shim.py, or the real peft package)
        for k in [k for k in sd if ".base_layer." in k]:
            sd.setdefault(k.replace(".base_layer.", "."), sd[k])
        return sd

    def _check_ready(self):
        w = self.lm_head.weight
        if not w.is_cuda or w.dtype != bf16:
            raise _lib.MplError("medplib_b200 runs in bf16 on a CUDA device only: call model.to(torch.bfloat16).cuda() "
                                "(the reference's --precision bf16 path); there is no CPU / fp32 fallback")
        _lib.load()

    def _llama(self):
        if "llama" not in self._eng:
            self._check_ready()
            params = self._params()
            adapted = [(n, mod) for n, mod in self.named_modules() if hasattr(mod, "lora_A") and hasattr(mod, "scaling")]
            if adapted:
                # validation in the middle of training (train_ds_medplib.py validate() runs the peft-wrapped model):
                # the inference kernels read plain weights, so the engine is built on MERGED copies W + s B A of the
                # adapted matrices (offline plumbing, rebuilt after every optimizer step); the parameters stay untouched
                self._merged = {}
                with torch.no_grad():
                    for n, mod in adapted:
                        A, Bm = mod.lora_A["default"].weight, mod.lora_B["default"].weight
                        w = (mod.weight.float() + float(mod.scaling["default"]) * (Bm.float() @ A.float())).to(mod.weight.dtype)
                        self._merged[n + ".weight"] = w
                params = dict(params)
                params.update(self._merged)
            self._eng["llama"] = engine.LlamaEngine(params, llama_dims(self.config), "model.")
        return self._eng["llama"]

    def _sam_encoder(self):
        if "sam_enc" not in self._eng:
            self._check_ready()
            enc = self.model.visual_model.image_encoder
            self._eng["sam_enc"] = engine.SamEncoderEngine(self._params(), enc.cfg, "model.visual_model.image_encoder.")
        return self._eng["sam_enc"]

    def _mask_decoder(self):
        if "sam_dec" not in self._eng:
            self._check_ready()
            grid = self.model.visual_model.prompt_encoder.image_embedding_size[0]
            self._eng["sam_dec"] = engine.MaskDecoderEngine(self._params(), "model.visual_model.", grid)
        return self._eng["sam_dec"]

    # ------------------------------------------------------------------ vision / glue
    def encode_images(self, images, region_flag=False, region_geo_sampler=False):
        """medplib_arch.py:198-212 -> (raw CLIP features, projected [compressed] features, region feature map)."""
        m = self.get_model()
        if region_geo_sampler and not hasattr(m, "region_geo_sampler"):
            raise _lib.MplError("region_geo_sampler requested but the model was built without config.region_geo_sampler "
                                "(the reference keeps the module commented out, medplib_arch.py:134-142)")
        feats = m.get_vision_tower()(images)
        proj = m.mm_projector
        if isinstance(proj, nn.Linear):
            x = ops.linear(feats, proj.weight, bias=proj.bias)
        else:
            x = ops.linear(feats, proj[0].weight, bias=proj[0].bias, act="gelu")
            x = ops.linear(x, proj[2].weight, bias=proj[2].bias)
        if getattr(self.config, "mm_token_compress", False):
            c = m.mm_token_compressor
            x = ops.pool_layernorm(x, c.norm.weight, c.norm.bias, c.num_tokens, c.norm.eps)
            x = ops.linear(x, c.proj.weight, bias=c.proj.bias)
        rmap = None
        if region_flag:
            # :204-208 — the geometric sampler reads the raw CLIP features, the default path the adapted ones
            rmap = feats if region_geo_sampler else ops.linear(feats, m.region_fea_adapter.weight,
                                                               bias=m.region_fea_adapter.bias)
        return feats, x, rmap

    def encode_masks(self, mask_images):
        """MaskTokenEncoder.forward (medplib_arch.py:95-108): 4x conv3x3 s2 + GELU as im2col GEMMs, adaptive pool fused
        with nothing to normalise (identity LN weights are not assumed: pool, Linear, LayerNorm)."""
        me = self.get_model().mask_encoder
        x = mask_images
        if x.dim() == 3:
            x = x.unsqueeze(1)
        x = x[:, :1].to(bf16).permute(0, 2, 3, 1).contiguous()  # NHWC, C=1 -> pad channels to 8 for 16-byte rows
        n = x.shape[0]
        x = torch.cat([x, x.new_zeros(*x.shape[:3], 7)], dim=-1)
        for i in (0, 2, 4, 6):
            conv = me.encoder[i]
            H = x.shape[1]
            Cin = x.shape[-1]
            w = conv.weight
            if w.shape[1] != Cin:  # first conv: zero-pad the input-channel dim to match
                w = torch.cat([w, w.new_zeros(w.shape[0], Cin - w.shape[1], 3, 3)], dim=1)
            w2 = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()
            cols = ops.im2col_nhwc(x, 3, 3, 2, 1)
            Ho = (H + 2 - 3) // 2 + 1
            x = ops.linear(cols, w2, bias=conv.bias, act="gelu").view(n, Ho, Ho, -1)
        feats = x.reshape(n, -1, x.shape[-1])  # [n, 441, 256] == features.flatten(2).transpose(1, 2)
        ones = torch.ones(feats.shape[-1], dtype=bf16, device=feats.device)
        pooled = _adaptive_pool_tokens(feats, me.num_tokens)
        y = ops.linear(pooled, me.proj.weight, bias=me.proj.bias)
        del ones
        return ops.layernorm(y, me.norm.weight, me.norm.bias, me.norm.eps)

    def extract_region_feature(self, region_feature_map, region_masks, original_dtype=None, return_dtype=None,
                               raw_feature_map=None, raw_out=None):
        """medplib_arch.py:580-614. Coordinates are rounded to the run dtype exactly like the reference does.
        raw_feature_map / raw_out (train step): the same points also sample the raw CLIP features — the region feature
        is linear in the adapter (sample-mean and Linear commute), so d adapter = d feature (x) sampled raw feature."""
        out = []
        assert len(region_feature_map) == len(region_masks), f"{len(region_feature_map)}, {len(region_masks)}"
        for bi, (fmap, masks) in enumerate(zip(region_feature_map, region_masks)):
            if len(masks) == 0:
                out.append(None)
                continue
            h = w = int(round(fmap.shape[0] ** 0.5))
            feats = []
            for m in masks:
                hw = torch.tensor([m.shape[0], m.shape[1]], device=m.device)[None]
                nz = m.nonzero() / hw
                if nz.shape[0] > self.get_model().max_sample_point:
                    nz = nz[torch.randperm(nz.shape[0])[: self.get_model().max_sample_point]]
                pts = nz.flip(dims=(1,)).to(fmap.dtype).float()
                feats.append(ops.region_sample_mean(fmap, pts.to(fmap.device), h, w))
                if raw_out is not None:
                    raw_out.append(ops.region_sample_mean(raw_feature_map[bi], pts.to(fmap.device), h, w))
            out.append(torch.stack(feats))
        return out

    def prepare_inputs_labels_for_multimodal(self, input_ids, attention_mask, past_key_values, labels, images,
                                             region_masks, valid_region_masks_bool, mask_images=None,
                                             image_token_types=None):
        """medplib_arch.py:217-527. The sentinel splice is planned on the host (like the reference's Python loops) and
        executed as ONE row-gather kernel over [embed_tokens ; image / mask / region feature rows]."""
        region_flag = region_masks is not None and len(region_masks) > 0
        vt = self.get_vision_tower()
        if vt is None or images is None or input_ids.shape[1] == 1:
            if past_key_values is not None and vt is not None and images is not None and input_ids.shape[1] == 1:
                attention_mask = torch.ones((attention_mask.shape[0], past_key_values.len + 1),
                                            dtype=attention_mask.dtype, device=attention_mask.device)
            return input_ids, attention_mask, past_key_values, None, labels
        per_token = False
        feat_blocks = []  # list of [n_tokens, D] tensors in sentinel order
        block_kinds = None  # ICL separate mode: "image" / "mask" per block (None: all image blocks)
        mask_cat = None

        if image_token_types is not None and mask_images is not None and len(mask_images) > 0:
            assert type(images) is list or images.ndim == 5
            assert not region_flag
            raw_f, img_f, _ = self.encode_images(torch.cat([im for im in images], dim=0))
            mask_cat = torch.cat([m for m in mask_images], dim=0)
            msk_f = self.encode_masks(mask_cat)
            ii = mi = 0
            block_kinds = []
            for types_ in image_token_types:
                for t in types_:
                    if t == "mask":
                        feat_blocks.append(msk_f[mi]); mi += 1
                        block_kinds.append("mask")
                    else:
                        feat_blocks.append(img_f[ii]); ii += 1
                        block_kinds.append("image")
            per_token = True
        elif type(images) is list or images.ndim == 5:
            assert not region_flag
            raw_f, img_f, _ = self.encode_images(torch.cat([im for im in images], dim=0))
            feat_blocks = [f for f in img_f]
            per_token = True
        else:
            geo = region_flag and bool(getattr(self.config, "region_geo_sampler", False))  # medplib_arch.py:229
            raw_f, img_f, rmap = self.encode_images(images, region_flag, geo)
            feat_blocks = [f for f in img_f]
        region_features = None
        valid = None
        self._region_ctx = None
        raw_samples = None
        if region_flag:
            valid = [any(item) for item in valid_region_masks_bool]
            vsel = torch.tensor(valid, device=rmap.device)
            rmap = rmap[vsel]
            adapter = self.get_model().region_fea_adapter
            want_grad = (labels is not None and torch.is_grad_enabled() and self.training
                         and adapter.weight.requires_grad and not geo)
            raw_samples = [] if want_grad else None
            if geo:  # medplib_arch.py:285-289
                region_features = self.get_model().region_geo_sampler(rmap, region_masks, original_dtype=raw_f.dtype,
                                                                      return_dtype=img_f.dtype)
            else:
                region_features = self.extract_region_feature(rmap, region_masks,
                                                              raw_feature_map=raw_f[vsel] if want_grad else None,
                                                              raw_out=raw_samples)

        # ---- host-side plan: one int per output row
        D = self.config.hidden_size
        use_se = getattr(self.config, "mm_use_im_start_end", False)
        ids_host = input_ids.tolist()
        lab_host = labels.tolist() if labels is not None else None
        offsets, off = [], 0
        for f in feat_blocks:
            offsets.append(off)
            off += f.shape[0]
        extra_rows = []  # region vectors appended after the image/mask blocks
        plans, new_labels = [], ([] if labels is not None else None)
        img_i = 0
        for b, ids in enumerate(ids_host):
            plan, lab = [], []
            cur_lab = lab_host[b] if lab_host is not None else None
            if IMAGE_TOKEN_INDEX not in ids:
                plans.append(list(ids))
                if new_labels is not None:
                    new_labels.append(list(cur_lab))
                if not per_token:
                    img_i += 1
                continue
            pos = 0
            while IMAGE_TOKEN_INDEX in ids[pos:]:
                s = ids.index(IMAGE_TOKEN_INDEX, pos)
                if region_flag:
                    assert REGION_TOKEN_INDEX not in ids[pos:s]
                plan.extend(ids[pos:s])
                n_f = feat_blocks[img_i].shape[0]
                plan.extend(-(offsets[img_i] + r) - 2 for r in range(n_f))
                if cur_lab is not None:
                    lab.extend(cur_lab[pos:s]); lab.extend([IGNORE_INDEX] * n_f)
                if use_se:
                    plan.extend(ids[s + 1:s + 2])
                    if cur_lab is not None:
                        lab.extend(cur_lab[s + 1:s + 2])
                    pos = s + 2
                else:
                    pos = s + 1
                img_i += 1
            rest = ids[pos:]
            if len(rest) > 0:
                ridx = [i for i, t in enumerate(rest) if t == REGION_TOKEN_INDEX]
                text = [t for t in rest if t != REGION_TOKEN_INDEX]
                if cur_lab is not None:
                    lab.extend(cur_lab[pos:])
                if region_flag and valid[b]:
                    k = sum(valid[:b + 1]) - 1
                    for j, at in enumerate(ridx):
                        text.insert(at, -(off + len(extra_rows)) - 2)
                        extra_rows.append(region_features[k][j])
                plan.extend(text)
            plans.append(plan)
            if new_labels is not None:
                new_labels.append(lab)
        T = max(len(p) for p in plans)
        ragged = any(len(p) != T for p in plans)
        dev = input_ids.device
        idx = torch.tensor([p + [-1] * (T - len(p)) for p in plans], dtype=torch.int32).to(dev)
        feats_all = torch.cat(feat_blocks + ([torch.stack(extra_rows)] if extra_rows else []), dim=0).contiguous()
        embeds = ops.gather_rows(idx.reshape(-1), self.model.embed_tokens.weight, feats_all, D=D).view(len(plans), T, D)
        self._splice_idx = idx.reshape(-1)  # adjoint of the splice (embed_tokens gradient) in the train step
        # ---- train step: what the backward of the vision-side adapters needs (mm_projector: scripts/train_stage2.sh;
        # mm_token_compressor / mask_encoder: scripts/train_medplib_icl.sh). Feature row k of a block sits at output row
        # pos[k] of the spliced prompt (-1: the prompt never used it).
        self._proj_ctx = None
        mdl = self.get_model()
        proj = mdl.mm_projector
        compress = bool(getattr(self.config, "mm_token_compress", False))
        tracking = labels is not None and torch.is_grad_enabled() and self.training
        proj_tr = tracking and any(p.requires_grad for p in proj.parameters())
        comp_tr = tracking and compress and any(p.requires_grad for p in mdl.mm_token_compressor.parameters())
        menc_tr = tracking and mask_cat is not None and any(p.requires_grad for p in mdl.mask_encoder.parameters())
        if proj_tr and isinstance(proj, nn.Linear):
            self._proj_ctx = dict(unsupported="mm_projector gradients are built for the mlp2x_gelu projector "
                                  "(multimodal_projector/builder.py:39-46); freeze a linear projector")
        elif proj_tr or comp_tr or menc_tr:
            n_rows = off
            pos = [-1] * n_rows
            for b, p in enumerate(plans):
                for t, v in enumerate(p):
                    if v <= -2 and -v - 2 < n_rows:
                        pos[-v - 2] = b * T + t
            img_pos, mask_pos = [], []
            for bi, f in enumerate(feat_blocks):
                seg = pos[offsets[bi]:offsets[bi] + f.shape[0]]
                (mask_pos if block_kinds is not None and block_kinds[bi] == "mask" else img_pos).extend(seg)
            raw_all = raw_f if not isinstance(raw_f, list) else torch.cat(raw_f, 0)
            self._proj_ctx = dict(pos=torch.tensor(img_pos, dtype=torch.int32, device=dev),
                                  feats=raw_all.reshape(-1, raw_all.shape[-1]), n_img=int(raw_all.shape[0]),
                                  compress=compress, proj_train=proj_tr, comp_train=comp_tr)
            if menc_tr:
                self._proj_ctx["mask"] = dict(pos=torch.tensor(mask_pos, dtype=torch.int32, device=dev), images=mask_cat)
        if raw_samples:
            # output rows that hold region features, in extra_rows order (plan entries <= -(off) - 2)
            pos = [b * T + t for b, p in enumerate(plans) for t, v in enumerate(p) if v <= -off - 2]
            order = sorted(range(len(pos)), key=lambda i: -(plans[pos[i] // T][pos[i] % T]) - 2 - off)
            pos = [pos[i] for i in order]
            assert len(pos) == len(raw_samples)
            self._region_ctx = dict(pos=torch.tensor(pos, dtype=torch.int32, device=dev),
                                    sampled=torch.stack(raw_samples).contiguous())
        lab_t = None
        if labels is not None:
            lab_t = torch.tensor([l + [IGNORE_INDEX] * (T - len(l)) for l in new_labels], dtype=labels.dtype, device=dev)
        if attention_mask is not None:
            n_in = input_ids.shape[1]
            rows = []
            for b, p in enumerate(plans):
                n_b = len(p) if ragged else T
                left = torch.ones(n_b - n_in, dtype=attention_mask.dtype, device=dev)
                right = torch.zeros(T - n_b, dtype=attention_mask.dtype, device=dev)
                rows.append(torch.cat((left, attention_mask[b], right)))
            attention_mask = torch.stack(rows)
        return None, attention_mask, past_key_values, embeds, lab_t

    # ------------------------------------------------------------------ causal-LM forward (MedPLIBMoELlamaForCausalLM.forward)
    def _lm_forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                    inputs_embeds=None, labels=None, use_cache=None, output_attentions=None,
                    output_hidden_states=None, images=None, return_dict=None, region_masks=None,
                    valid_region_masks_bool=None, mask_images=None, image_token_types=None, logits_rows="all",
                    **_unused):
        """medplib_moe_llama.py:324-438. past_key_values is a medplib_b200.engine.KVCache (or None)."""
        if output_attentions:
            raise _lib.MplError("attention maps are never materialised by the fused attention kernels")
        self._splice_idx = None
        self._region_ctx = None
        if inputs_embeds is None:
            input_ids, attention_mask, past_key_values, inputs_embeds, labels = \
                self.prepare_inputs_labels_for_multimodal(input_ids, attention_mask, past_key_values, labels, images,
                                                          region_masks, valid_region_masks_bool,
                                                          mask_images=mask_images, image_token_types=image_token_types)
        if inputs_embeds is None:
            idx = input_ids.reshape(-1).to(torch.int32)
            inputs_embeds = ops.gather_rows(idx, self.model.embed_tokens.weight).view(*input_ids.shape, -1)
            self._splice_idx = idx
        if labels is not None:
            return self._lm_forward_train(inputs_embeds, attention_mask, labels, _unused.get("moe_noise"))
        eng = self._llama()
        B, T, D = inputs_embeds.shape
        cache = past_key_values
        if cache is None:
            cache = eng.new_cache(B, T + (int(use_cache or 0) and 64))
        elif cache.len + T > cache.Tmax:
            cache = _grow_cache(eng, cache, cache.len + T + 256)
        kv_mask = None
        if attention_mask is not None and not bool(attention_mask.all()):
            kv_mask = attention_mask
        x = inputs_embeds.to(bf16).contiguous().clone()
        out = eng.forward(x, cache, training=self.training, kv_mask=kv_mask,
                          want_hidden_states=bool(output_hidden_states), want_router=True)
        hidden = out["last_hidden_state"]
        self._fire_gate_hooks(out["gate_logits"])
        if logits_rows == "last":
            logits = ops.linear(hidden[:, -1], self.lm_head.weight, out_dtype=torch.float32).unsqueeze(1)
        else:
            logits = ops.linear(hidden, self.lm_head.weight, out_dtype=torch.float32)
        # one entry per MoE layer, as MoELlamaModel_forward collects them (medplib_moe_llama.py:265-283): dense layers add none
        moe_losses = [out["l_aux"][li] for li in self._moe_layer_ids()] if out["l_aux"] is not None else []
        moe_loss = self.router_aux_loss_coef * sum(moe_losses) if moe_losses else None
        loss = None
        hs = out["hidden_states"] if output_hidden_states else None
        return MoECausalLMOutputWithPast(loss=loss, moe_loss=moe_loss, logits=logits,
                                         past_key_values=cache if use_cache else None,
                                         hidden_states=hs if hs is not None else (hidden,),
                                         moe_loss_list=moe_losses)

    # ------------------------------------------------------------------ train step (medplib_b200/train.py)
    def trainer(self, **hyper):
        """The Trainer (gradient arena + optimizer + tape nodes) bound to the current set of trainable parameters.
        Rebuilt when that set changes (attach_lora / requires_grad edits) or when hyper-parameters are passed."""
        from .. import train as _train
        tr = getattr(self, "_trainer", None)
        if tr is None or hyper or tr._sig != _train.Trainer.signature(self):
            self._trainer = tr = _train.Trainer(self, **hyper)
        return tr

    def refresh_trained(self):
        """After an optimizer step: engines that hold REPACKED / MERGED copies of trainable weights (mask decoder,
        LoRA-merged decoder matrices) are rebuilt on their next use."""
        self._eng.pop("sam_dec", None)
        if getattr(self, "_merged", None):
            self._eng.pop("llama", None)
            self._merged = None

    def _lm_forward_train(self, inputs_embeds, attention_mask, labels, moe_noise=None):
        """Training branch of medplib_moe_llama.py:324-438: activations kept, loss = shifted CE + coef * sum(l_aux),
        differentiable through medplib_b200.train's tape nodes (loss.backward() fills the gradient arena)."""
        tr = self.trainer()
        pc = getattr(self, "_proj_ctx", None)
        if pc is not None and "unsupported" in pc:
            raise _lib.MplError(pc["unsupported"])
        kv_mask = None
        if attention_mask is not None and not bool(attention_mask.all()):
            kv_mask = attention_mask
        hidden, l_aux = tr.stack_hidden(inputs_embeds, kv_mask=kv_mask, moe_noise=moe_noise,
                                        splice_idx=self._splice_idx, region_ctx=getattr(self, "_region_ctx", None),
                                        proj_ctx=getattr(self, "_proj_ctx", None))
        self._region_ctx = self._proj_ctx = None
        self._fire_gate_hooks(tr.last_gate_logits)
        loss, logits = tr.head_ce(hidden, labels)
        moe_losses = list(l_aux.unbind(0)) if l_aux.numel() > 0 else []
        moe_loss = None
        if moe_losses:
            moe_loss = self.router_aux_loss_coef * l_aux.sum()
            loss = loss + moe_loss
        return MoECausalLMOutputWithPast(loss=loss, moe_loss=moe_loss, logits=logits, past_key_values=None,
                                         hidden_states=(hidden,), moe_loss_list=moe_losses)

    def _fire_gate_hooks(self, gate_logits):
        """vqa_infer.py:157-165 registers forward hooks on the `wg` Linears to read the router logits.
        gate_logits: the engine's [L, S * Emax] f32 buffer indexed by TRANSFORMER layer (rows of layer l packed [S, E_l]),
        or the train path's list with one [S, E_l] tensor per MoE layer."""
        if gate_logits is None:
            return
        per_moe = isinstance(gate_logits, (list, tuple))
        mi = 0
        for li, layer in enumerate(self.model.layers):
            if not isinstance(layer.mlp, M.MoE):
                continue
            wg = layer.mlp.deepspeed_moe.gate.wg
            if wg._forward_hooks:
                if per_moe:
                    lg = gate_logits[mi]
                else:
                    E = wg.weight.shape[0]
                    flat = gate_logits[li].reshape(-1)
                    lg = flat[:(flat.numel() // gate_logits.shape[-1]) * E].view(-1, E)
                for hook in list(wg._forward_hooks.values()):
                    hook(wg, (None,), lg)
            mi += 1

    def _moe_layer_ids(self):
        return [li for li, layer in enumerate(self.model.layers) if isinstance(layer.mlp, M.MoE)]

    def forward(self, **kwargs):
        if "past_key_values" in kwargs:
            return self._lm_forward(**kwargs)
        return self.model_forward(**kwargs)

    # ------------------------------------------------------------------ generation
    @torch.no_grad()
    def generate(self, input_ids=None, images=None, attention_mask=None, region_masks=None,
                 valid_region_masks_bool=None, mask_images=None, image_token_types=None, do_sample=False,
                 temperature=1.0, top_p=None, num_beams=1, max_new_tokens=512, use_cache=True,
                 output_hidden_states=False, return_dict_in_generate=False, eos_token_id=None,
                 forced_tokens=None, output_scores=False, **_unused):
        """Greedy / sampled decoding with the reference's keyword surface (vqa_infer.py:430-442, MedPLIB.py:592-606).
        Prefill runs the tcgen05 path over the spliced prompt; every decode step is rmsnorm -> streaming GEMMs ->
        KV-cache attention -> MoE scatter/gather over B rows, fed by the device-side argmax (no host sync per token
        except the periodic EOS check). forced_tokens {step: id} overrides the choice at that step (benchmarks)."""
        if num_beams != 1:
            raise _lib.MplError("beam search is not built (the reference always decodes with num_beams=1)")
        eos = eos_token_id if eos_token_id is not None else getattr(self.config, "eos_token_id", None)
        eos = None if (isinstance(eos, int) and eos < 0) else eos
        dev = input_ids.device
        B = input_ids.shape[0]
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids, dtype=torch.bool)
        _, am, _, embeds, _ = self.prepare_inputs_labels_for_multimodal(
            input_ids, attention_mask, None, None, images, region_masks, valid_region_masks_bool,
            mask_images=mask_images, image_token_types=image_token_types)
        if embeds is None:
            embeds = ops.gather_rows(input_ids.reshape(-1).to(torch.int32), self.model.embed_tokens.weight) \
                .view(*input_ids.shape, -1)
            am = attention_mask
        eng = self._llama()
        T = embeds.shape[1]
        D = embeds.shape[2]
        Tmax = T + max_new_tokens
        cache = eng.new_cache(B, Tmax)
        kv_mask = None
        if not bool(am.all()):
            kv_mask = torch.ones((B, Tmax), dtype=torch.bool, device=dev)
            kv_mask[:, :T] = am
        hidden_all = torch.empty((B, Tmax - 1, D), dtype=bf16, device=dev) if output_hidden_states else None
        out = eng.forward(embeds.clone(), cache, kv_mask=kv_mask, want_router=True)
        self._fire_gate_hooks(out["gate_logits"])
        h = out["last_hidden_state"]
        if hidden_all is not None:
            hidden_all[:, :T] = h
        last = h[:, -1].contiguous()
        tokens = torch.zeros((B, max_new_tokens), dtype=torch.int64, device=dev)
        logits = torch.empty((B, self.lm_head.weight.shape[0]), dtype=torch.float32, device=dev)
        scores = [] if output_scores else None
        for step in range(max_new_tokens):
            ops.linear(last, self.lm_head.weight, out_dtype=torch.float32, out=logits)
            if scores is not None:
                scores.append(logits.clone())
            if do_sample and temperature and temperature > 0:
                nxt = _sample(logits, temperature, top_p)
            else:
                nxt = ops.argmax(logits)
            if forced_tokens is not None and step in forced_tokens:
                nxt = torch.full_like(nxt, forced_tokens[step])
            tokens[:, step] = nxt
            if step == max_new_tokens - 1:
                break
            if eos is not None and (step % 16 == 15):
                hit = (tokens[:, :step + 1] == eos).any(dim=1)
                if bool(hit.all()):
                    break
            x = ops.gather_rows(nxt.to(torch.int32), self.model.embed_tokens.weight).view(B, 1, D)
            out = eng.forward(x, cache, kv_mask=kv_mask, want_router=bool(self._has_gate_hooks()))
            self._fire_gate_hooks(out["gate_logits"])
            last = out["last_hidden_state"][:, 0]
            if hidden_all is not None:
                hidden_all[:, T + step] = last
        tokens_h = tokens
        n_new = max_new_tokens
        if eos is not None:
            is_eos = tokens == eos
            if bool(is_eos.any()):
                first = torch.where(is_eos.any(1), is_eos.float().argmax(1), torch.full((B,), max_new_tokens - 1,
                                                                                       device=dev))
                n_new = int(first.max().item()) + 1
                pad = getattr(self.config, "pad_token_id", None)
                pad = eos if pad is None else pad
                ar = torch.arange(max_new_tokens, device=dev)[None]
                tokens_h = torch.where(ar > first[:, None], torch.full_like(tokens, pad), tokens)
        seq = torch.cat([input_ids, tokens_h[:, :n_new]], dim=1)
        if not return_dict_in_generate:
            return seq
        lh = hidden_all[:, :T + n_new - 1] if hidden_all is not None else None
        return GenerateOutput(sequences=seq, hidden_states=None, last_hidden_state=lh,
                              scores=tuple(scores) if scores is not None else None)

    def _has_gate_hooks(self):
        for layer in self.model.layers:
            if isinstance(layer.mlp, M.MoE) and len(layer.mlp.deepspeed_moe.gate.wg._forward_hooks) > 0:
                return True
        return False

    # ------------------------------------------------------------------ SAM side
    def get_visual_embs(self, pixel_values):
        """MedPLIB.py:274-285 -> [B, 256, g, g] (the reference loops over images; here the batch runs together)."""
        with torch.no_grad():
            tok = self._sam_encoder().forward(pixel_values)  # [B, g*g, C] token-major
            B, T, C = tok.shape
            g = int(round(T ** 0.5))
            return tok.view(B, g, g, C).permute(0, 3, 1, 2)

    def is_empty_tensor(self, tensor):
        return tensor.numel() == 0

    def expand_embedding(self, image_embeddings, valid_mask_bool):
        """MedPLIB.py:292-308: repeat each image's embedding once per valid mask."""
        if valid_mask_bool is None or len(valid_mask_bool) == 0:
            return image_embeddings
        reps = [len(m) if m else 0 for m in valid_mask_bool]
        idx = [i for i, r in enumerate(reps) for _ in range(r)]
        return image_embeddings[torch.tensor(idx, device=image_embeddings.device, dtype=torch.long)]

    def build_seg_token_mask(self, input_ids, image_token_len=None, image_token_lengths=None):
        """MedPLIB.py:310-355 (position t is marked when token t+1 is <SEG>; IMAGE sentinels expand to their length)."""
        if image_token_len is None:
            image_token_len = (getattr(self.config, "mm_compressed_token_count", 256)
                               if getattr(self.config, "mm_token_compress", False)
                               else self.get_model().get_vision_tower().num_patches)
        rows = []
        ids_host = input_ids.tolist()
        for b, ids in enumerate(ids_host):
            cur, k = [], 0
            for t, tok in enumerate(ids):
                if tok == IMAGE_TOKEN_INDEX:
                    n = image_token_len
                    if image_token_lengths is not None and len(image_token_lengths) > b and \
                            len(image_token_lengths[b]) > k:
                        n = image_token_lengths[b][k]
                    k += 1
                    cur.extend([False] * n)
                else:
                    cur.append(t + 1 < len(ids) and ids[t + 1] == self.seg_token_idx)
            rows.append(cur)
        T = max(len(r) for r in rows)
        return torch.tensor([r + [False] * (T - len(r)) for r in rows], dtype=torch.bool, device=input_ids.device)

    def _seg_embeddings(self, hidden_rows):
        """text_hidden_fcs on the selected rows only (the reference runs it on every position and gathers after)."""
        fc = self.model.text_hidden_fcs[0]
        h = ops.linear(hidden_rows.contiguous(), fc[0].weight, bias=fc[0].bias, act="relu")
        return ops.linear(h, fc[2].weight, bias=fc[2].bias)

    def _decode_masks(self, pred_embeddings, image_embeddings, resize_list, size_list):
        """Prompt encoder (text) + mask decoder + postprocess for every [SEG] embedding (MedPLIB.py:473-502)."""
        dec = self._mask_decoder()
        B, C, g, _ = image_embeddings.shape
        tok = image_embeddings.permute(0, 2, 3, 1).reshape(B, g * g, C)
        pred_masks, pred_ious = [], []
        for i in range(len(pred_embeddings)):
            low, iou = dec.forward(tok[i], pred_embeddings[i])
            pm = self.postprocess_masks(low, input_size=resize_list[i], original_size=size_list[i])
            pred_masks.append(pm[:, 0])
            pred_ious.append(iou[:, 0])
        return pred_masks, pred_ious

    def postprocess_masks(self, masks, input_size, original_size):
        """MedPLIB.py:682-701 (crop by mask_size - input_size — a no-op at 64x64 — then bilinear resize)."""
        if masks.dim() == 3:
            masks = masks.unsqueeze(0)
        pad_h = masks.shape[-2] - input_size[0]
        pad_w = masks.shape[-1] - input_size[1]
        top, left = pad_h // 2, pad_w // 2
        oh, ow = masks.shape[-2] - pad_h, masks.shape[-1] - pad_w
        masks = masks[:, :, top:top + oh, left:left + ow]
        n, c = masks.shape[:2]
        out = ops.bilinear_resize(masks.reshape(n * c, masks.shape[-2], masks.shape[-1]), tuple(original_size))
        return out.view(n, c, *out.shape[-2:])

    # ------------------------------------------------------------------ model_forward / evaluate
    def model_forward(self, images, images_clip, input_ids, region_masks=None, labels=None, attention_mask=None,
                      offset=None, masks_list=None, label_list=None, resize_list=None, inference=False,
                      seg_flag=True, valid_mask_bool=None, rp_flag=False, valid_region_masks_bool=None, **kwargs):
        """MedPLIB.py:364-572. inference=True is the single-pass grounding forward ([SEG] in the prompt)."""
        if not inference:
            return self._model_forward_train(images, images_clip, input_ids, region_masks, labels, attention_mask,
                                             masks_list, label_list, resize_list, seg_flag, valid_mask_bool,
                                             valid_region_masks_bool, **kwargs)
        with torch.no_grad():
            if seg_flag:
                image_embeddings = self.expand_embedding(self.get_visual_embs(images), valid_mask_bool)
                seg_token_mask = self.build_seg_token_mask(input_ids,
                                                           image_token_lengths=kwargs.get("image_token_lengths"))
            out = self._lm_forward(images=images_clip, attention_mask=attention_mask, input_ids=input_ids,
                                   labels=None, output_hidden_states=False, region_masks=region_masks,
                                   valid_region_masks_bool=valid_region_masks_bool,
                                   mask_images=kwargs.get("mask_images"),
                                   image_token_types=kwargs.get("image_token_types"), logits_rows="last")
            if not seg_flag:
                return {"pred_masks": [], "gt_masks": masks_list}
            last_hidden = out.hidden_states[-1]
            seg_token_mask = seg_token_mask[:, :last_hidden.shape[1]]
            pred_embeddings = self._seg_embeddings(last_hidden[seg_token_mask])
            if kwargs.get("icl_image_counts") is not None and masks_list is not None and len(masks_list) > 0:
                pred_embeddings = pred_embeddings[-len(masks_list):]
            sizes = [tuple(l.shape) for l in label_list]
            pred_masks, _ = self._decode_masks(pred_embeddings, image_embeddings, resize_list, sizes)
            return {"pred_masks": pred_masks, "gt_masks": masks_list}

    def _model_forward_train(self, images, images_clip, input_ids, region_masks, labels, attention_mask, masks_list,
                             label_list, resize_list, seg_flag, valid_mask_bool, valid_region_masks_bool, **kwargs):
        """MedPLIB.py:403-572 with inference=False: CE (+ router aux) loss, and with seg_flag the four mask losses of
        every [SEG] row; returns the reference's 10-key dict. out["loss"].backward() fills trainer().arena."""
        if seg_flag:
            image_embeddings = self.expand_embedding(self.get_visual_embs(images), valid_mask_bool)
            seg_token_mask = self.build_seg_token_mask(input_ids, image_token_lengths=kwargs.get("image_token_lengths"))
        out = self._lm_forward(images=images_clip, attention_mask=attention_mask, input_ids=input_ids, labels=labels,
                               region_masks=region_masks, valid_region_masks_bool=valid_region_masks_bool,
                               mask_images=kwargs.get("mask_images"),
                               image_token_types=kwargs.get("image_token_types"), moe_noise=kwargs.get("moe_noise"))
        ce_loss = out.loss * self.ce_loss_weight
        if not seg_flag:
            z = torch.zeros_like(ce_loss)
            return {"loss": ce_loss, "ce_loss": ce_loss, "mask_bce_loss": z, "mask_dice_loss": z, "mask_loss": z,
                    "unscale_mask_bce_loss": z, "unscale_mask_dice_loss": z, "unscale_mask_loss": z,
                    "unscale_mask_iou_loss": z, "unscale_mask_focal_loss": z}
        from .. import mask_train
        hidden = out.hidden_states[-1]
        seg_token_mask = seg_token_mask[:, :hidden.shape[1]]
        tr = self.trainer()
        rows = mask_train.select_rows(hidden, seg_token_mask)
        pred_embeddings = mask_train.text_hidden_fcs(tr, self.model.text_hidden_fcs[0], rows)
        if kwargs.get("icl_image_counts") is not None and len(masks_list) > 0:
            pred_embeddings = pred_embeddings[-len(masks_list):]
        sizes = [tuple(l.shape) for l in label_list]
        sums = mask_train.mask_head_losses(tr, self, pred_embeddings, image_embeddings, resize_list, sizes, masks_list)
        num_masks = sums["num_masks"]
        un_bce = sums["bce"] / (num_masks + 1e-8)
        un_dice = sums["dice"] / (num_masks + 1e-8)
        un_iou = sums["iou"] / (num_masks + 1e-8)
        un_focal = sums["focal"] / (num_masks + 1e-8)
        bce, dice = self.bce_loss_weight * un_bce, self.dice_loss_weight * un_dice
        iou, focal = self.iou_loss_weight * un_iou, self.focal_loss_weight * un_focal
        mask_loss = bce + dice + iou + focal
        return {"loss": ce_loss + mask_loss, "ce_loss": ce_loss, "mask_bce_loss": bce, "mask_dice_loss": dice,
                "mask_loss": mask_loss, "unscale_mask_bce_loss": un_bce, "unscale_mask_dice_loss": un_dice,
                "unscale_mask_loss": un_bce + un_dice + un_iou + un_focal, "unscale_mask_iou_loss": un_iou,
                "unscale_mask_focal_loss": un_focal}

    def evaluate(self, images_clip, images, input_ids, resize_list, original_size_list, region_masks=[],
                 valid_region_masks_bool=[], max_new_tokens=512, tokenizer=None, attention_mask=None,
                 inference_demo=False, mask_images=None, image_token_types=None, image_token_lengths=None,
                 forced_tokens=None):
        """MedPLIB.py:574-680: generate, take the hidden state in front of the first <SEG> (or position -2 when there is
        none), project it, run SAM-Med2D on `images` and decode one mask."""
        with torch.no_grad():
            # The SAM-Med2D image encoder does not depend on the language model (the reference runs it after decoding,
            # MedPLIB.py:648): it is enqueued on a side stream so its small launch-bound GEMMs fill the SMs that the
            # M = 615 prefill tiles and the HBM-bound decode steps leave idle; joined before the mask decoder.
            side = image_embeddings = None
            if getattr(self, "overlap_vision", True) and images.is_cuda:
                main = torch.cuda.current_stream()
                side = self._side_stream = getattr(self, "_side_stream", None) or torch.cuda.Stream(device=images.device)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    image_embeddings = self.get_visual_embs(images)
            gen = self.generate(images=images_clip, input_ids=input_ids, region_masks=region_masks,
                                valid_region_masks_bool=valid_region_masks_bool, mask_images=mask_images,
                                image_token_types=image_token_types, max_new_tokens=max_new_tokens, do_sample=False,
                                num_beams=1, output_hidden_states=True, return_dict_in_generate=True, use_cache=True,
                                attention_mask=attention_mask, forced_tokens=forced_tokens)
            output_ids = gen.sequences
            hidden = gen.last_hidden_state
            # ONE device -> host copy of the generated ids (the number of <SEG> rows decides shapes, so the host has to
            # see them once); the <SEG> bookkeeping then runs on the host copy and the selected hidden row is a view --
            # no `.any()` reduction, no boolean-index kernels and no second / third synchronisation on the device
            ids_host = output_ids.cpu()
            has_seg = bool((ids_host[:, 1:] == self.seg_token_idx).any())
            if not has_seg and inference_demo:
                if side is not None:
                    torch.cuda.current_stream().wait_stream(side)
                return output_ids, []
            seg_mask = self.build_seg_token_mask(ids_host, image_token_lengths=image_token_lengths)
            seg_mask = seg_mask[:, :hidden.shape[1]]  # App. B-1: the last mask entry is always False
            pos = seg_mask.nonzero()  # row-major order == the order of hidden[seg_mask]; the reference keeps the first
            if pos.shape[0] > 0:
                rows = hidden[int(pos[0, 0]), int(pos[0, 1])][None]
            else:
                rows = hidden[:1, -2:-1, :].squeeze(1)
            pred_embeddings = self._seg_embeddings(rows)
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)
                image_embeddings.record_stream(torch.cuda.current_stream())
            else:
                image_embeddings = self.get_visual_embs(images)
            sizes = [tuple(o.shape) for o in original_size_list]
            pred_masks, _ = self._decode_masks(pred_embeddings, image_embeddings, resize_list, sizes)
        return output_ids, pred_masks


def _adaptive_pool_tokens(x, n_out):
    """AdaptiveAvgPool1d over the token axis: windows [floor(i*T/n), ceil((i+1)*T/n)), fp32 mean (mpl_token_pool)."""
    from .. import train_ops
    return train_ops.token_pool(x, n_out)


def _grow_cache(eng, cache, new_tmax):
    new = eng.new_cache(cache.B, new_tmax)
    new.k[:, :, :, :cache.len] = cache.k[:, :, :, :cache.len]
    new.v[:, :, :, :cache.len] = cache.v[:, :, :, :cache.len]
    new.len = cache.len
    return new


def _sample(logits, temperature, top_p):
    probs = torch.softmax(logits / temperature, dim=-1)
    if top_p is not None and top_p < 1.0:
        sp, si = torch.sort(probs, descending=True, dim=-1)
        keep = (sp.cumsum(-1) - sp) < top_p
        sp = sp * keep
        probs = torch.zeros_like(probs).scatter_(1, si, sp)
        probs = probs / probs.sum(-1, keepdim=True)
    return torch.multinomial(probs, 1).squeeze(1)
