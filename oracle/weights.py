"""Synthetic random-init state dicts with the reference's parameter names and shapes (TEST INFRASTRUCTURE ONLY — see
oracle/__init__.py). Shapes follow SURVEY.md Appendix C; names follow model/MedPLIB.py, medplib_arch.py,
medplib_moe_llama.py:617-635 (expert keys), HF CLIPVisionModel and model/segment_anything_med2d/build_sam.py."""
import torch


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _n(g, *shape, std=0.02):
    return torch.randn(*shape, generator=g) * std


def llama(cfg, seed=0, dtype=torch.bfloat16, prefix="model.", wg_std=0.5):
    """cfg: hidden_size, intermediate_size, num_layers, num_heads, vocab_size, moe{num_experts or None}."""
    g = _g(seed)
    D, F, L, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_layers"], cfg["vocab_size"]
    E = (cfg.get("moe") or {}).get("num_experts")
    sd = {prefix + "embed_tokens.weight": _n(g, V, D).to(dtype), "lm_head.weight": _n(g, V, D).to(dtype),
          prefix + "norm.weight": (1 + _n(g, D, std=0.1)).to(dtype)}
    s = D ** -0.5
    for i in range(L):
        p = f"{prefix}layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[f"{p}self_attn.{n}.weight"] = _n(g, D, D, std=s).to(dtype)
        sd[p + "input_layernorm.weight"] = (1 + _n(g, D, std=0.1)).to(dtype)
        sd[p + "post_attention_layernorm.weight"] = (1 + _n(g, D, std=0.1)).to(dtype)
        if E:
            sd[p + "mlp.deepspeed_moe.gate.wg.weight"] = _n(g, E, D, std=wg_std).float()
            for e in range(E):
                ep = f"{p}mlp.deepspeed_moe.experts.deepspeed_experts.{e}."
                sd[ep + "gate_proj.weight"] = _n(g, F, D, std=s).to(dtype)
                sd[ep + "up_proj.weight"] = _n(g, F, D, std=s).to(dtype)
                sd[ep + "down_proj.weight"] = _n(g, D, F, std=F ** -0.5).to(dtype)
        else:
            sd[p + "mlp.gate_proj.weight"] = _n(g, F, D, std=s).to(dtype)
            sd[p + "mlp.up_proj.weight"] = _n(g, F, D, std=s).to(dtype)
            sd[p + "mlp.down_proj.weight"] = _n(g, D, F, std=F ** -0.5).to(dtype)
    return sd


def clip(cfg, seed=1, dtype=torch.bfloat16, prefix=""):
    """cfg: hidden_size, intermediate_size, num_layers, num_heads, image_size, patch_size."""
    g = _g(seed)
    D, M, L, P = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_layers"], cfg["patch_size"]
    n = (cfg["image_size"] // P) ** 2 + 1
    p = prefix + "vision_model."
    sd = {p + "embeddings.class_embedding": _n(g, D, std=0.5).to(dtype),
          p + "embeddings.patch_embedding.weight": _n(g, D, 3, P, P, std=(3 * P * P) ** -0.5).to(dtype),
          p + "embeddings.position_embedding.weight": _n(g, n, D, std=0.5).to(dtype)}
    for nm in ("pre_layrnorm", "post_layernorm"):
        sd[f"{p}{nm}.weight"] = (1 + _n(g, D, std=0.1)).to(dtype)
        sd[f"{p}{nm}.bias"] = _n(g, D, std=0.1).to(dtype)
    for i in range(L):
        lp = f"{p}encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[f"{lp}self_attn.{nm}.weight"] = _n(g, D, D, std=D ** -0.5).to(dtype)
            sd[f"{lp}self_attn.{nm}.bias"] = _n(g, D, std=0.1).to(dtype)
        for nm in ("layer_norm1", "layer_norm2"):
            sd[f"{lp}{nm}.weight"] = (1 + _n(g, D, std=0.1)).to(dtype)
            sd[f"{lp}{nm}.bias"] = _n(g, D, std=0.1).to(dtype)
        sd[lp + "mlp.fc1.weight"] = _n(g, M, D, std=D ** -0.5).to(dtype)
        sd[lp + "mlp.fc1.bias"] = _n(g, M, std=0.1).to(dtype)
        sd[lp + "mlp.fc2.weight"] = _n(g, D, M, std=M ** -0.5).to(dtype)
        sd[lp + "mlp.fc2.bias"] = _n(g, D, std=0.1).to(dtype)
    return sd


def sam_encoder(cfg, seed=2, dtype=torch.bfloat16, prefix="", adapter=True):
    """cfg: embed_dim, depth, num_heads, image_size, patch_size, out_chans. Windows 14, global blocks {2,5,8,11}."""
    g = _g(seed)
    D, depth, H, P, O = cfg["embed_dim"], cfg["depth"], cfg["num_heads"], cfg["patch_size"], cfg["out_chans"]
    grid = cfg["image_size"] // P
    hd = D // H
    p = prefix
    sd = {p + "patch_embed.proj.weight": _n(g, D, 3, P, P, std=(3 * P * P) ** -0.5).to(dtype),
          p + "patch_embed.proj.bias": _n(g, D, std=0.1).to(dtype),
          p + "pos_embed": _n(g, 1, grid, grid, D, std=0.5).to(dtype)}
    for i in range(depth):
        bp = f"{p}blocks.{i}."
        s = grid if i in (2, 5, 8, 11) else 14
        for nm in ("norm1", "norm2"):
            sd[f"{bp}{nm}.weight"] = (1 + _n(g, D, std=0.1)).to(dtype)
            sd[f"{bp}{nm}.bias"] = _n(g, D, std=0.1).to(dtype)
        sd[bp + "attn.qkv.weight"] = _n(g, 3 * D, D, std=D ** -0.5).to(dtype)
        sd[bp + "attn.qkv.bias"] = _n(g, 3 * D, std=0.1).to(dtype)
        sd[bp + "attn.proj.weight"] = _n(g, D, D, std=D ** -0.5).to(dtype)
        sd[bp + "attn.proj.bias"] = _n(g, D, std=0.1).to(dtype)
        sd[bp + "attn.rel_pos_h"] = _n(g, 2 * s - 1, hd, std=0.2).to(dtype)
        sd[bp + "attn.rel_pos_w"] = _n(g, 2 * s - 1, hd, std=0.2).to(dtype)
        sd[bp + "mlp.lin1.weight"] = _n(g, 4 * D, D, std=D ** -0.5).to(dtype)
        sd[bp + "mlp.lin1.bias"] = _n(g, 4 * D, std=0.1).to(dtype)
        sd[bp + "mlp.lin2.weight"] = _n(g, D, 4 * D, std=(4 * D) ** -0.5).to(dtype)
        sd[bp + "mlp.lin2.bias"] = _n(g, D, std=0.1).to(dtype)
        if adapter:
            sd[bp + "Adapter.channel.0.weight"] = _n(g, D // 4, D, std=D ** -0.5).to(dtype)
            sd[bp + "Adapter.channel.2.weight"] = _n(g, D, D // 4, std=(D // 4) ** -0.5).to(dtype)
            sd[bp + "Adapter.spatial.0.weight"] = _n(g, D, D, 3, 3, std=(9 * D) ** -0.5).to(dtype)
            sd[bp + "Adapter.spatial.2.weight"] = _n(g, D, D, 4, 4, std=(4 * D) ** -0.5).to(dtype)
            sd[bp + "Adapter.norm.weight"] = (1 + _n(g, D, std=0.1)).to(dtype)
            sd[bp + "Adapter.norm.bias"] = _n(g, D, std=0.1).to(dtype)
    sd[p + "neck.0.weight"] = _n(g, O, D, 1, 1, std=D ** -0.5).to(dtype)
    sd[p + "neck.2.weight"] = _n(g, O, O, 3, 3, std=(9 * O) ** -0.5).to(dtype)
    for nm in ("neck.1", "neck.3"):
        sd[f"{p}{nm}.weight"] = (1 + _n(g, O, std=0.1)).to(dtype)
        sd[f"{p}{nm}.bias"] = _n(g, O, std=0.1).to(dtype)
    return sd


def sam_head(seed=3, dtype=torch.bfloat16, prefix="", dim=256, mlp=2048, depth=2, n_mask=4):
    """prompt_encoder (text path) + mask_decoder of build_sam_vit_b."""
    g = _g(seed)
    pe, p = prefix + "prompt_encoder.", prefix + "mask_decoder."
    sd = {pe + "pe_layer.positional_encoding_gaussian_matrix": torch.randn(2, dim // 2, generator=g).to(dtype),
          pe + "no_mask_embed.weight": _n(g, 1, dim, std=0.5).to(dtype),
          p + "iou_token.weight": _n(g, 1, dim, std=0.5).to(dtype),
          p + "mask_tokens.weight": _n(g, n_mask, dim, std=0.5).to(dtype)}

    def attn(ap, internal):
        for nm in ("q_proj", "k_proj", "v_proj"):
            sd[f"{ap}{nm}.weight"] = _n(g, internal, dim, std=dim ** -0.5).to(dtype)
            sd[f"{ap}{nm}.bias"] = _n(g, internal, std=0.1).to(dtype)
        sd[ap + "out_proj.weight"] = _n(g, dim, internal, std=internal ** -0.5).to(dtype)
        sd[ap + "out_proj.bias"] = _n(g, dim, std=0.1).to(dtype)

    def ln(k, d):
        sd[k + ".weight"] = (1 + _n(g, d, std=0.1)).to(dtype)
        sd[k + ".bias"] = _n(g, d, std=0.1).to(dtype)

    for i in range(depth):
        lp = f"{p}transformer.layers.{i}."
        attn(lp + "self_attn.", dim)
        attn(lp + "cross_attn_token_to_image.", dim // 2)
        attn(lp + "cross_attn_image_to_token.", dim // 2)
        for j in (1, 2, 3, 4):
            ln(f"{lp}norm{j}", dim)
        sd[lp + "mlp.lin1.weight"] = _n(g, mlp, dim, std=dim ** -0.5).to(dtype)
        sd[lp + "mlp.lin1.bias"] = _n(g, mlp, std=0.1).to(dtype)
        sd[lp + "mlp.lin2.weight"] = _n(g, dim, mlp, std=mlp ** -0.5).to(dtype)
        sd[lp + "mlp.lin2.bias"] = _n(g, dim, std=0.1).to(dtype)
    attn(p + "transformer.final_attn_token_to_image.", dim // 2)
    ln(p + "transformer.norm_final_attn", dim)
    sd[p + "output_upscaling.0.weight"] = _n(g, dim, dim // 4, 2, 2, std=dim ** -0.5).to(dtype)
    sd[p + "output_upscaling.0.bias"] = _n(g, dim // 4, std=0.1).to(dtype)
    ln(p + "output_upscaling.1", dim // 4)
    sd[p + "output_upscaling.3.weight"] = _n(g, dim // 4, dim // 8, 2, 2, std=(dim // 4) ** -0.5).to(dtype)
    sd[p + "output_upscaling.3.bias"] = _n(g, dim // 8, std=0.1).to(dtype)

    def mlp3(mp, out):
        dims = [dim, dim, dim, out]
        for j in range(3):
            sd[f"{mp}layers.{j}.weight"] = _n(g, dims[j + 1], dims[j], std=dims[j] ** -0.5).to(dtype)
            sd[f"{mp}layers.{j}.bias"] = _n(g, dims[j + 1], std=0.1).to(dtype)

    for i in range(n_mask):
        mlp3(f"{p}output_hypernetworks_mlps.{i}.", dim // 8)
    mlp3(p + "iou_prediction_head.", n_mask)
    return sd
