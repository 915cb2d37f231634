"""Stage-by-stage op-level replay of the CLIP / SAM-encoder block sequences against oracle intermediates (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from medplib_b200 import ops, engine
from oracle import clip, sam, weights
bf16 = torch.bfloat16
dev = "cuda"

def rep(name, got, ref):
    got, ref = got.float().cpu().reshape(ref.shape), ref.float()
    err = (got - ref).abs().max().item(); sc = ref.abs().max().item()
    bad = (~torch.isfinite(got)).sum().item()
    print(f"{name:28s} err={err:.4e} scale={sc:.3e} rel={err/max(sc,1e-9):.3e} nonfinite={bad}", flush=True)

def clip_debug(cfg, B):
    sd = weights.clip(cfg, seed=23); p = "vision_model."
    img = torch.randn(B, 3, cfg["image_size"], cfg["image_size"], generator=torch.Generator().manual_seed(2)).to(bf16)
    D, P, H = cfg["hidden_size"], cfg["patch_size"], cfg["num_heads"]; hd = D // H
    g = {k: v.to(dev) for k, v in sd.items()}
    # oracle intermediates
    x_ref = clip.embeddings(sd, p, img, cfg)
    k = 3 * P * P; kp = (k + 7) // 8 * 8
    pw = torch.zeros(D, kp, dtype=bf16, device=dev); pw[:, :k] = g[p + "embeddings.patch_embedding.weight"].reshape(D, k)
    cols = ops.im2col_patch(img.to(dev), P, kp)
    patch = ops.linear(cols, pw)
    rep("patch conv", patch, F.conv2d(img, sd[p + "embeddings.patch_embedding.weight"], stride=P).flatten(2).transpose(1, 2))
    x = ops.clip_embed(patch, g[p + "embeddings.class_embedding"], g[p + "embeddings.position_embedding.weight"], B)
    rep("embeddings", x, x_ref)
    x = ops.layernorm(x, g[p + "pre_layrnorm.weight"], g[p + "pre_layrnorm.bias"], 1e-5)
    xr = F.layer_norm(x_ref, (D,), sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], 1e-5)
    rep("pre_ln", x, xr)
    lp = p + "encoder.layers.0."
    h = ops.layernorm(x, g[lp + "layer_norm1.weight"], g[lp + "layer_norm1.bias"], 1e-5)
    hr = F.layer_norm(xr, (D,), sd[lp + "layer_norm1.weight"], sd[lp + "layer_norm1.bias"], 1e-5)
    rep("ln1", h, hr)
    T = x.shape[1]
    qkv = torch.empty(B * T, 3 * D, dtype=bf16, device=dev)
    ops.linear(h, [g[lp + f"self_attn.{n}_proj.weight"] for n in "qkv"], bias=[g[lp + f"self_attn.{n}_proj.bias"] for n in "qkv"],
               out=[qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]])
    qr = F.linear(hr, sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.q_proj.bias"])
    kr = F.linear(hr, sd[lp + "self_attn.k_proj.weight"], sd[lp + "self_attn.k_proj.bias"])
    vr = F.linear(hr, sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"])
    rep("q (fused buf)", qkv[:, :D], qr.reshape(-1, D)); rep("k (fused buf)", qkv[:, D:2 * D], kr.reshape(-1, D)); rep("v (fused buf)", qkv[:, 2 * D:], vr.reshape(-1, D))
    q4 = qkv.view(B, T, 3, H, hd)
    o = ops.attention(q4[:, :, 0], q4[:, :, 1], q4[:, :, 2], hd ** -0.5)
    ar = clip.attention({k2: v for k2, v in sd.items()}, lp + "self_attn.", hr, cfg)
    # reference attention before out_proj
    qf, kf, vf = (t.view(B, T, H, hd).transpose(1, 2).float() for t in (qr, kr, vr))
    pr = torch.softmax(qf @ kf.transpose(-1, -2) * hd ** -0.5, -1) @ vf
    rep("attention (fused in)", o, pr.transpose(1, 2))
    o2 = ops.attention(q4[:, :, 0].contiguous(), q4[:, :, 1].contiguous(), q4[:, :, 2].contiguous(), hd ** -0.5)
    rep("attention (contig in)", o2, pr.transpose(1, 2))
    x2 = x.clone()
    ops.linear(o.reshape(B * T, D), g[lp + "self_attn.out_proj.weight"], bias=g[lp + "self_attn.out_proj.bias"], residual=x2.view(-1, D), out=x2.view(-1, D))
    x2r = xr + ar
    rep("out_proj + res (in place)", x2, x2r)
    h2 = ops.layernorm(x2, g[lp + "layer_norm2.weight"], g[lp + "layer_norm2.bias"], 1e-5)
    h2r = F.layer_norm(x2r, (D,), sd[lp + "layer_norm2.weight"], sd[lp + "layer_norm2.bias"], 1e-5)
    m1 = ops.linear(h2, g[lp + "mlp.fc1.weight"], bias=g[lp + "mlp.fc1.bias"], act="quick_gelu")
    m1r = clip.quick_gelu(F.linear(h2r, sd[lp + "mlp.fc1.weight"], sd[lp + "mlp.fc1.bias"]))
    rep("fc1 quick_gelu", m1, m1r)
    x3 = x2.clone()
    ops.linear(m1, g[lp + "mlp.fc2.weight"], bias=g[lp + "mlp.fc2.bias"], residual=x3.view(-1, D), out=x3.view(-1, D))
    rep("fc2 + res", x3, x2r + F.linear(m1r, sd[lp + "mlp.fc2.weight"], sd[lp + "mlp.fc2.bias"]))
    ecfg = dict(cfg)
    eng = engine.ClipEngine(g, ecfg, "", select_layer=1)
    rep("ENGINE 1 layer", eng.forward(img.to(dev)), clip.vision_tower(sd, "", img, cfg, select_layer=1))
    eng0 = engine.ClipEngine(g, ecfg, "", select_layer=0)
    rep("ENGINE 0 layers", eng0.forward(img.to(dev)), clip.vision_tower(sd, "", img, cfg, select_layer=0))

if __name__ == "__main__":
    torch.manual_seed(0)
    clip_debug(dict(hidden_size=128, intermediate_size=256, num_layers=3, num_heads=2, image_size=56, patch_size=14), 2)
    clip_debug(dict(hidden_size=1024, intermediate_size=4096, num_layers=3, num_heads=16, image_size=336, patch_size=14), 1)
