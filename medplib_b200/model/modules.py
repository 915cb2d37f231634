"""Parameter-holding nn.Modules with the reference's module paths and parameter names.

The arithmetic does NOT run through these modules' ``forward``: MedPLIBForCausalLM hands their parameters (in place) to
the native stack runners of libmedplib_b200.so (medplib_b200/engine.py). The modules exist so that everything that
walks the module tree keeps working unchanged against the new build (SURVEY.md §8b "naming invariants"): LoRA target
matching over ``nn.Linear`` names (train_ds_medplib.py:265-291), ``--sft_modules`` substring matching (:316-326), the
``wg`` forward hooks of model/eval/vqa_infer.py:157-165, and state-dict loading of the published checkpoints
(expert keys ``...mlp.deepspeed_moe.experts.deepspeed_experts.{e}...``, medplib_moe_llama.py:617-635).
"""
import copy

import torch
import torch.nn as nn

from .. import _lib


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise _lib.MplError(f"{type(self).__name__} holds parameters for the fused sm_100a stack; it is not called "
                            "directly (there is no eager fallback)")


# ------------------------------------------------------------------------------------------------------------ LLaMA
class LlamaRMSNorm(_Holder):
    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps


class LlamaAttention(_Holder):
    def __init__(self, config):
        super().__init__()
        D = config.hidden_size
        self.q_proj = nn.Linear(D, D, bias=False)
        self.k_proj = nn.Linear(D, D, bias=False)
        self.v_proj = nn.Linear(D, D, bias=False)
        self.o_proj = nn.Linear(D, D, bias=False)


class LlamaMLP(_Holder):
    def __init__(self, config):
        super().__init__()
        D, F = config.hidden_size, config.intermediate_size
        self.gate_proj = nn.Linear(D, F, bias=False)
        self.up_proj = nn.Linear(D, F, bias=False)
        self.down_proj = nn.Linear(F, D, bias=False)


class TopKGate(_Holder):
    """deepspeed.moe.sharded_moe.TopKGate: the gate Linear is kept in fp32."""

    def __init__(self, model_dim, num_experts, k, capacity_factor, eval_capacity_factor, min_capacity):
        super().__init__()
        self.wg = nn.Linear(model_dim, num_experts, bias=False).float()
        self.k, self.capacity_factor, self.eval_capacity_factor = k, capacity_factor, eval_capacity_factor
        self.min_capacity = min_capacity

    def _apply(self, fn, *a, **k):
        # model.to(bf16) must not down-cast the gate (DeepSpeed re-casts it to fp32 in TopKGate.forward)
        super()._apply(fn, *a, **k)
        self.wg.float()
        return self


class Experts(_Holder):
    def __init__(self, expert, num_local_experts):
        super().__init__()
        self.deepspeed_experts = nn.ModuleList([copy.deepcopy(expert) for _ in range(num_local_experts)])
        self.num_local_experts = num_local_experts
        for e in self.deepspeed_experts:
            for p in e.parameters():  # attributes DeepSpeed sets and the train driver's optimizer grouping reads
                p.allreduce = False
                p.group_name = "ep_size_1"


class MOELayer(_Holder):
    def __init__(self, gate, experts):
        super().__init__()
        self.gate = gate
        self.experts = experts


class MoE(_Holder):
    """Same constructor and module layout as deepspeed.moe.layer.MoE (model/MedPLIB.py:253-263)."""

    def __init__(self, hidden_size, expert, num_experts=1, ep_size=1, k=1, capacity_factor=1.0,
                 eval_capacity_factor=1.0, min_capacity=4, use_residual=False, **_):
        super().__init__()
        if use_residual:
            raise _lib.MplError("use_residual=True (Residual-MoE) is not used by MedPLIB and is not implemented")
        if ep_size not in (1, None):
            raise _lib.MplError("expert parallelism is degenerate in the reference (ep_size=1); only ep_size=1 is built")
        if k not in (1, 2):
            raise _lib.MplError("DeepSpeed TopKGate supports k in {1, 2}")
        self.num_experts, self.ep_size, self.use_residual = num_experts, 1, False
        self.deepspeed_moe = MOELayer(
            TopKGate(hidden_size, num_experts, k, capacity_factor, eval_capacity_factor, min_capacity),
            Experts(expert, num_experts))


class LlamaDecoderLayer(_Holder):
    def __init__(self, config):
        super().__init__()
        self.self_attn = LlamaAttention(config)
        self.mlp = LlamaMLP(config)
        self.input_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)


# ------------------------------------------------------------------------------------------------------------ CLIP
class _CLIPEmbeddings(_Holder):
    def __init__(self, c):
        super().__init__()
        self.class_embedding = nn.Parameter(torch.randn(c.hidden_size))
        self.patch_embedding = nn.Conv2d(3, c.hidden_size, c.patch_size, stride=c.patch_size, bias=False)
        n = (c.image_size // c.patch_size) ** 2 + 1
        self.position_embedding = nn.Embedding(n, c.hidden_size)
        self.register_buffer("position_ids", torch.arange(n).expand((1, -1)), persistent=False)


class _CLIPAttention(_Holder):
    def __init__(self, c):
        super().__init__()
        D = c.hidden_size
        self.k_proj, self.v_proj = nn.Linear(D, D), nn.Linear(D, D)
        self.q_proj, self.out_proj = nn.Linear(D, D), nn.Linear(D, D)


class _CLIPMLP(_Holder):
    def __init__(self, c):
        super().__init__()
        self.fc1 = nn.Linear(c.hidden_size, c.intermediate_size)
        self.fc2 = nn.Linear(c.intermediate_size, c.hidden_size)


class _CLIPEncoderLayer(_Holder):
    def __init__(self, c):
        super().__init__()
        self.self_attn = _CLIPAttention(c)
        self.layer_norm1 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.mlp = _CLIPMLP(c)
        self.layer_norm2 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _CLIPEncoder(_Holder):
    def __init__(self, c):
        super().__init__()
        self.layers = nn.ModuleList([_CLIPEncoderLayer(c) for _ in range(c.num_hidden_layers)])


class _CLIPVisionTransformer(_Holder):
    def __init__(self, c):
        super().__init__()
        self.embeddings = _CLIPEmbeddings(c)
        self.pre_layrnorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.encoder = _CLIPEncoder(c)
        self.post_layernorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class CLIPVisionModelHolder(_Holder):
    """Same parameter names as transformers.CLIPVisionModel (vision_model.*)."""

    def __init__(self, c):
        super().__init__()
        self.config = c
        self.vision_model = _CLIPVisionTransformer(c)


# ------------------------------------------------------------------------------------------------------------ SAM-Med2D
class _LayerNorm2d(_Holder):
    def __init__(self, n, eps=1e-6):
        super().__init__()
        self.weight, self.bias, self.eps = nn.Parameter(torch.ones(n)), nn.Parameter(torch.zeros(n)), eps


class _AdapterLayer(_Holder):
    def __init__(self, D):
        super().__init__()
        self.norm = nn.LayerNorm(D)
        self.channel = nn.Sequential(nn.Linear(D, D // 4, bias=False), nn.ReLU(), nn.Linear(D // 4, D, bias=False),
                                     nn.Sigmoid())
        self.spatial = nn.Sequential(nn.Conv2d(D, D, 3, stride=2, padding=1, bias=False), nn.ReLU(),
                                     nn.ConvTranspose2d(D, D, 4, stride=2, padding=1, bias=False), nn.ReLU())


class _SamEncAttention(_Holder):
    def __init__(self, D, heads, size):
        super().__init__()
        self.qkv, self.proj = nn.Linear(D, 3 * D), nn.Linear(D, D)
        self.rel_pos_h = nn.Parameter(torch.zeros(2 * size - 1, D // heads))
        self.rel_pos_w = nn.Parameter(torch.zeros(2 * size - 1, D // heads))


class _MLPBlock(_Holder):
    def __init__(self, D, M):
        super().__init__()
        self.lin1, self.lin2 = nn.Linear(D, M), nn.Linear(M, D)


class _SamBlock(_Holder):
    def __init__(self, D, heads, size, adapter):
        super().__init__()
        self.norm1 = nn.LayerNorm(D, eps=1e-6)
        self.attn = _SamEncAttention(D, heads, size)
        self.norm2 = nn.LayerNorm(D, eps=1e-6)
        self.mlp = _MLPBlock(D, 4 * D)
        if adapter:
            self.Adapter = _AdapterLayer(D)


class _PatchEmbed(_Holder):
    def __init__(self, D, P):
        super().__init__()
        self.proj = nn.Conv2d(3, D, P, stride=P)


class ImageEncoderViT(_Holder):
    def __init__(self, img_size=256, patch_size=16, embed_dim=768, depth=12, num_heads=12, out_chans=256,
                 window_size=14, global_attn_indexes=(2, 5, 8, 11), adapter=True):
        super().__init__()
        self.img_size = img_size
        g = img_size // patch_size
        self.patch_embed = _PatchEmbed(embed_dim, patch_size)
        self.pos_embed = nn.Parameter(torch.zeros(1, g, g, embed_dim))
        self.blocks = nn.ModuleList([
            _SamBlock(embed_dim, num_heads, g if i in global_attn_indexes else window_size, adapter)
            for i in range(depth)])
        self.neck = nn.Sequential(nn.Conv2d(embed_dim, out_chans, 1, bias=False), _LayerNorm2d(out_chans),
                                  nn.Conv2d(out_chans, out_chans, 3, padding=1, bias=False), _LayerNorm2d(out_chans))
        self.cfg = dict(embed_dim=embed_dim, depth=depth, num_heads=num_heads, image_size=img_size,
                        patch_size=patch_size, out_chans=out_chans)


class _PositionEmbeddingRandom(_Holder):
    def __init__(self, n):
        super().__init__()
        self.register_buffer("positional_encoding_gaussian_matrix", torch.randn((2, n)))


class PromptEncoder(_Holder):
    def __init__(self, embed_dim=256, grid=16, mask_in_chans=16):
        super().__init__()
        self.embed_dim, self.image_embedding_size = embed_dim, (grid, grid)
        self.pe_layer = _PositionEmbeddingRandom(embed_dim // 2)
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(4)])
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, 2, stride=2), _LayerNorm2d(mask_in_chans // 4), nn.GELU(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, 2, stride=2), _LayerNorm2d(mask_in_chans), nn.GELU(),
            nn.Conv2d(mask_in_chans, embed_dim, 1))
        self.no_mask_embed = nn.Embedding(1, embed_dim)


class _SamAttention(_Holder):
    def __init__(self, D, downsample=1):
        super().__init__()
        I = D // downsample
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = nn.Linear(D, I), nn.Linear(D, I), nn.Linear(D, I), nn.Linear(I, D)


class _TwoWayBlock(_Holder):
    def __init__(self, D, M):
        super().__init__()
        self.self_attn = _SamAttention(D)
        self.norm1 = nn.LayerNorm(D)
        self.cross_attn_token_to_image = _SamAttention(D, 2)
        self.norm2 = nn.LayerNorm(D)
        self.mlp = _MLPBlock(D, M)
        self.norm3 = nn.LayerNorm(D)
        self.norm4 = nn.LayerNorm(D)
        self.cross_attn_image_to_token = _SamAttention(D, 2)


class _TwoWayTransformer(_Holder):
    def __init__(self, D=256, depth=2, M=2048):
        super().__init__()
        self.layers = nn.ModuleList([_TwoWayBlock(D, M) for _ in range(depth)])
        self.final_attn_token_to_image = _SamAttention(D, 2)
        self.norm_final_attn = nn.LayerNorm(D)


class _MLP(_Holder):
    def __init__(self, i, h, o, n):
        super().__init__()
        dims = [i] + [h] * (n - 1) + [o]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))


class MaskDecoder(_Holder):
    def __init__(self, D=256, n_multimask=3):
        super().__init__()
        self.transformer = _TwoWayTransformer(D)
        self.iou_token = nn.Embedding(1, D)
        self.num_mask_tokens = n_multimask + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, D)
        self.output_upscaling = nn.Sequential(nn.ConvTranspose2d(D, D // 4, 2, stride=2), _LayerNorm2d(D // 4),
                                              nn.GELU(), nn.ConvTranspose2d(D // 4, D // 8, 2, stride=2), nn.GELU())
        self.output_hypernetworks_mlps = nn.ModuleList([_MLP(D, D, D // 8, 3) for _ in range(self.num_mask_tokens)])
        self.iou_prediction_head = _MLP(D, 256, self.num_mask_tokens, 3)


class Sam(_Holder):
    """build_sam_vit_b (model/segment_anything_med2d/build_sam.py:51-61): image_size 256, adapters on."""

    def __init__(self, image_size=256, embed_dim=768, depth=12, num_heads=12):
        super().__init__()
        self.image_encoder = ImageEncoderViT(image_size, 16, embed_dim, depth, num_heads)
        self.prompt_encoder = PromptEncoder(256, image_size // 16)
        self.mask_decoder = MaskDecoder(256)
        self.register_buffer("pixel_mean", torch.tensor([123.675, 116.28, 103.53]).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor([58.395, 57.12, 57.375]).view(-1, 1, 1), False)


# ------------------------------------------------------------------------------------------------------------ glue
class TokenCompressor(_Holder):
    def __init__(self, hidden_size, num_tokens=256):
        super().__init__()
        self.num_tokens = num_tokens
        self.norm = nn.LayerNorm(hidden_size)
        self.proj = nn.Linear(hidden_size, hidden_size)


class MaskTokenEncoder(_Holder):
    def __init__(self, hidden_size, num_tokens=64):
        super().__init__()
        self.num_tokens = num_tokens
        self.encoder = nn.Sequential(nn.Conv2d(1, 64, 3, stride=2, padding=1), nn.GELU(),
                                     nn.Conv2d(64, 128, 3, stride=2, padding=1), nn.GELU(),
                                     nn.Conv2d(128, 256, 3, stride=2, padding=1), nn.GELU(),
                                     nn.Conv2d(256, 256, 3, stride=2, padding=1), nn.GELU())
        self.proj = nn.Linear(256, hidden_size)
        self.norm = nn.LayerNorm(hidden_size)
