"""Oracle: SAM-Med2D (ViT-B + per-block adapters, text-prompt encoder, two-way mask decoder)
(TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Functional restatement, over a state dict with the reference's parameter names, of
  model/segment_anything_med2d/modeling/image_encoder.py  (Adapter_Layer :18-56, ImageEncoderViT.forward :151-162,
      Block.forward :214-238, Attention.forward :280-296, window_partition/unpartition :299-345,
      get_rel_pos / add_decomposed_rel_pos :348-421, PatchEmbed :424-455)
  model/segment_anything_med2d/modeling/prompt_encoder.py (forward with text_embeds :140-187, get_dense_pe :62-71,
      PositionEmbeddingRandom :190-226)
  model/segment_anything_med2d/modeling/mask_decoder.py   (predict_masks :113-153, MLP :158-186)
  model/segment_anything_med2d/modeling/transformer.py    (TwoWayTransformer :62-106, TwoWayAttentionBlock :148-182,
      Attention.forward :213-244)
  model/segment_anything_med2d/modeling/common.py         (MLPBlock :13-26, LayerNorm2d :31-45)
with the hyper-parameters of build_sam_vit_b (build_sam.py:51-61,72-150): 12 blocks, dim 768, 12 heads, window 14,
global blocks {2,5,8,11}, patch 16, LayerNorm eps 1e-6 (adapter norm: default 1e-5), neck 256 channels.
Pinned against the reference's own modules (imported from /root/reference in the authoring container) through the
vectors in tests/golden/sam_*.pt (tests/golden/make_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

GLOBAL_ATTN = (2, 5, 8, 11)
WINDOW = 14


def layernorm2d(x, w, b, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def _rel_pos(q_size, k_size, rel_pos):
    max_rel = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != max_rel:
        r = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=max_rel, mode="linear")
        rel_pos = r.reshape(-1, max_rel).permute(1, 0)
    qc = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    rel = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rel_pos[rel.long()]


def encoder_attention(sd, p, x, num_heads):
    """Attention.forward with decomposed relative positions. x [B,H,W,C]."""
    B, H, W, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).reshape(B, H * W, 3, num_heads, -1)
    q, k, v = qkv.permute(2, 0, 3, 1, 4).reshape(3, B * num_heads, H * W, -1).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh = _rel_pos(H, H, sd[p + "rel_pos_h"])
    Rw = _rel_pos(W, W, sd[p + "rel_pos_w"])
    r_q = q.reshape(B * num_heads, H, W, hd).to(Rh.dtype)
    rel_h = torch.einsum("bhwc,hkc->bhwk", r_q, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", r_q, Rw)
    attn = (attn.view(-1, H, W, H, W) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(-1, H * W, H * W)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).view(B, num_heads, H, W, -1).permute(0, 2, 3, 1, 4).reshape(B, H, W, -1)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def window_partition(x, ws):
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)
    return x, (Hp, Wp)


def window_unpartition(win, ws, pad_hw, hw):
    Hp, Wp = pad_hw
    H, W = hw
    B = win.shape[0] // (Hp * Wp // ws // ws)
    x = win.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, -1)
    return x[:, :H, :W, :].contiguous()


def adapter(sd, p, x):
    """Adapter_Layer.forward: SE-style channel gate, conv3x3 s2 -> ReLU -> convT4x4 s2 -> ReLU, skip, LayerNorm."""
    x = x.permute(0, 3, 1, 2)
    B, C = x.shape[:2]
    pooled = x.mean((2, 3)) if x.dtype == torch.float32 else F.adaptive_avg_pool2d(x, 1).view(B, C)
    g = torch.sigmoid(F.linear(F.relu(F.linear(pooled, sd[p + "channel.0.weight"])), sd[p + "channel.2.weight"]))
    xc = g.view(B, C, 1, 1) * x
    s = F.relu(F.conv2d(xc, sd[p + "spatial.0.weight"], stride=2, padding=1))
    s = F.relu(F.conv_transpose2d(s, sd[p + "spatial.2.weight"], stride=2, padding=1))
    x = (x + s).permute(0, 2, 3, 1)
    return F.layer_norm(x, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)


def encoder_block(sd, p, x, num_heads, window):
    C = x.shape[-1]
    shortcut = x
    x = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    if window > 0:
        H, W = x.shape[1], x.shape[2]
        x, pad_hw = window_partition(x, window)
    x = encoder_attention(sd, p + "attn.", x, num_heads)
    if window > 0:
        x = window_unpartition(x, window, pad_hw, (H, W))
    x = shortcut + x
    xn = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    m = F.linear(F.gelu(F.linear(xn, sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"])),
                 sd[p + "mlp.lin2.weight"], sd[p + "mlp.lin2.bias"])
    if p + "Adapter.norm.weight" in sd:
        return x + m + adapter(sd, p + "Adapter.", xn)
    return x + m


def image_encoder(sd, prefix, images, num_heads=12, depth=None, patch=16):
    """ImageEncoderViT.forward. images [B,3,S,S] -> [B,256,S/16,S/16]. prefix e.g. 'model.visual_model.image_encoder.'"""
    p = prefix
    w = sd[p + "patch_embed.proj.weight"]
    x = F.conv2d(images.to(w.dtype), w, sd[p + "patch_embed.proj.bias"], stride=patch).permute(0, 2, 3, 1)
    x = x + sd[p + "pos_embed"]
    if depth is None:
        depth = 1 + max(int(k[len(p) + 7:].split(".")[0]) for k in sd if k.startswith(p + "blocks."))
    for i in range(depth):
        x = encoder_block(sd, f"{p}blocks.{i}.", x, num_heads, 0 if i in GLOBAL_ATTN else WINDOW)
    x = x.permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[p + "neck.0.weight"])
    x = layernorm2d(x, sd[p + "neck.1.weight"], sd[p + "neck.1.bias"])
    x = F.conv2d(x, sd[p + "neck.2.weight"], padding=1)
    return layernorm2d(x, sd[p + "neck.3.weight"], sd[p + "neck.3.bias"])


def dense_pe(sd, prefix, size):
    """PromptEncoder.get_dense_pe: fp32 [1, 2*F, h, w] random-Fourier positional encoding of the embedding grid."""
    G = sd[prefix + "pe_layer.positional_encoding_gaussian_matrix"].to(torch.float32)
    h, w = size
    grid = torch.ones((h, w), dtype=torch.float32)
    y = (grid.cumsum(0) - 0.5) / h
    x = (grid.cumsum(1) - 0.5) / w
    c = torch.stack([x, y], dim=-1)
    c = (2 * c - 1) @ G
    c = 2 * np.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1).permute(2, 0, 1).unsqueeze(0)


def prompt_encoder_text(sd, prefix, text_embeds, size):
    """PromptEncoder.forward(points=None, boxes=None, masks=None, text_embeds): sparse = cat(empty fp32, text) (promoted
    to fp32), dense = no_mask_embed broadcast over the grid."""
    bs = text_embeds.shape[0]
    sparse = torch.cat([torch.empty((bs, 0, text_embeds.shape[-1])), text_embeds], dim=1)
    dense = sd[prefix + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(bs, -1, size[0], size[1])
    return sparse, dense


def _mh_attention(sd, p, q, k, v, num_heads):
    """transformer.Attention.forward: inputs cast to the weight dtype, softmax(q k^T / sqrt(d)) v, out_proj."""
    wd = sd[p + "q_proj.weight"].dtype
    q = F.linear(q.to(wd), sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = F.linear(k.to(wd), sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(v.to(wd), sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])

    def sep(t):
        b, n, c = t.shape
        return t.reshape(b, n, num_heads, c // num_heads).transpose(1, 2)

    q, k, v = sep(q), sep(k), sep(v)
    attn = (q @ k.permute(0, 1, 3, 2)) / math.sqrt(q.shape[-1])
    attn = torch.softmax(attn, dim=-1)
    o = (attn @ v).transpose(1, 2)
    o = o.reshape(o.shape[0], o.shape[1], -1)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


def two_way_transformer(sd, p, image_embedding, image_pe, point_embedding, num_heads=8, depth=2):
    bs, c, h, w = image_embedding.shape
    keys = image_embedding.flatten(2).permute(0, 2, 1)
    key_pe = image_pe.flatten(2).permute(0, 2, 1)
    queries = point_embedding
    query_pe = point_embedding
    for i in range(depth):
        lp = f"{p}layers.{i}."
        if i == 0:
            queries = _mh_attention(sd, lp + "self_attn.", queries, queries, queries, num_heads)
        else:
            q = queries + query_pe
            queries = queries + _mh_attention(sd, lp + "self_attn.", q, q, queries, num_heads)
        queries = _ln(sd, lp + "norm1.", queries)
        q = queries + query_pe
        k = keys + key_pe
        queries = queries + _mh_attention(sd, lp + "cross_attn_token_to_image.", q, k, keys, num_heads)
        queries = _ln(sd, lp + "norm2.", queries)
        m = F.linear(F.relu(F.linear(queries, sd[lp + "mlp.lin1.weight"], sd[lp + "mlp.lin1.bias"])),
                     sd[lp + "mlp.lin2.weight"], sd[lp + "mlp.lin2.bias"])
        queries = _ln(sd, lp + "norm3.", queries + m)
        q = queries + query_pe
        k = keys + key_pe
        keys = keys + _mh_attention(sd, lp + "cross_attn_image_to_token.", k, q, queries, num_heads)
        keys = _ln(sd, lp + "norm4.", keys)
    q = queries + query_pe
    k = keys + key_pe
    queries = queries + _mh_attention(sd, p + "final_attn_token_to_image.", q, k, keys, num_heads)
    queries = _ln(sd, p + "norm_final_attn.", queries)
    return queries, keys


def _mlp(sd, p, x, n):
    for i in range(n):
        x = F.linear(x, sd[f"{p}layers.{i}.weight"], sd[f"{p}layers.{i}.bias"])
        if i < n - 1:
            x = F.relu(x)
    return x


def mask_decoder(sd, prefix, image_embeddings, image_pe, sparse, dense, multimask_output=False):
    """MaskDecoder.forward -> (masks [B,1|3,4h,4w], iou [B,1|3])."""
    p = prefix
    out_tok = torch.cat([sd[p + "iou_token.weight"], sd[p + "mask_tokens.weight"]], dim=0)
    n_mask = sd[p + "mask_tokens.weight"].shape[0]
    tokens = torch.cat((out_tok.unsqueeze(0).expand(sparse.size(0), -1, -1), sparse), dim=1)
    src = image_embeddings + dense
    pos_src = torch.repeat_interleave(image_pe, tokens.shape[0], dim=0)
    b, c, h, w = src.shape
    hs, src = two_way_transformer(sd, p + "transformer.", src, pos_src, tokens)
    iou_tok = hs[:, 0, :]
    mask_toks = hs[:, 1:1 + n_mask, :]
    src = src.transpose(1, 2).view(b, c, h, w)
    up = F.conv_transpose2d(src, sd[p + "output_upscaling.0.weight"], sd[p + "output_upscaling.0.bias"], stride=2)
    up = F.gelu(layernorm2d(up, sd[p + "output_upscaling.1.weight"], sd[p + "output_upscaling.1.bias"]))
    up = F.gelu(F.conv_transpose2d(up, sd[p + "output_upscaling.3.weight"], sd[p + "output_upscaling.3.bias"],
                                   stride=2))
    hyper = torch.stack([_mlp(sd, f"{p}output_hypernetworks_mlps.{i}.", mask_toks[:, i, :], 3)
                         for i in range(n_mask)], dim=1)
    b, c, h, w = up.shape
    masks = (hyper @ up.view(b, c, h * w)).view(b, -1, h, w)
    iou = _mlp(sd, p + "iou_prediction_head.", iou_tok, 3)
    sl = slice(1, None) if multimask_output else slice(0, 1)
    return masks[:, sl], iou[:, sl]
