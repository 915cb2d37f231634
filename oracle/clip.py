"""Oracle: CLIP ViT vision tower as the reference uses it (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Restates transformers==4.31.0 ``models/clip/modeling_clip.py`` {CLIPVisionEmbeddings, CLIPAttention, CLIPMLP,
CLIPEncoderLayer, CLIPVisionTransformer} (un-vendored; /root/reference/requirements.txt:137) and
the tower wrapper model/medplib/model/multimodal_encoder/clip_encoder.py:31-60 (hidden_states[select_layer], CLS
dropped). State-dict keys are HF's, relative to ``vision_model.`` (prefix given by the caller).
Pinned against the installed transformers-5.5 CLIPVisionModel in fp32 (tests/test_oracle_cpu.py).
"""
import torch
import torch.nn.functional as F


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def embeddings(sd, p, pixel_values, cfg):
    """CLIPVisionEmbeddings.forward: conv patch embed (no bias), prepend class token, add position embeddings."""
    w = sd[p + "embeddings.patch_embedding.weight"]
    x = F.conv2d(pixel_values.to(w.dtype), w, stride=cfg["patch_size"])
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1)
    return x + sd[p + "embeddings.position_embedding.weight"][None, : x.shape[1]]


def attention(sd, p, x, cfg):
    """CLIPAttention.forward (4.31): q scaled before the bmm; softmax in the input dtype."""
    B, T, D = x.shape
    H = cfg["num_heads"]
    hd = D // H
    scale = hd ** -0.5
    q = F.linear(x, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]) * scale
    k = F.linear(x, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(x, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    q = q.view(B, T, H, hd).transpose(1, 2)
    k = k.view(B, T, H, hd).transpose(1, 2)
    v = v.view(B, T, H, hd).transpose(1, 2)
    w = torch.matmul(q, k.transpose(-1, -2))
    w = F.softmax(w, dim=-1)
    o = torch.matmul(w, v).transpose(1, 2).reshape(B, T, D)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def encoder_layer(sd, p, x, cfg):
    eps = cfg.get("layer_norm_eps", 1e-5)
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)
    x = x + attention(sd, p + "self_attn.", h, cfg)
    h = F.layer_norm(x, (D,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)
    h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = quick_gelu(h)
    h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + h


def vision_tower(sd, prefix, images, cfg, select_layer=-2):
    """CLIPVisionTower.forward (clip_encoder.py:41-60): hidden_states[select_layer][:, 1:].

    hidden_states has num_layers+1 entries (post-pre_layrnorm embeddings + each layer's output); only the layers up to
    the selected one need to run (the reference runs all of them and discards the rest)."""
    p = prefix + "vision_model."
    eps = cfg.get("layer_norm_eps", 1e-5)
    x = embeddings(sd, p, images, cfg)
    x = F.layer_norm(x, (x.shape[-1],), sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], eps)
    L = cfg["num_layers"]
    n_run = select_layer if select_layer >= 0 else L + 1 + select_layer
    for i in range(n_run):
        x = encoder_layer(sd, f"{p}encoder.layers.{i}.", x, cfg)
    return x[:, 1:]
