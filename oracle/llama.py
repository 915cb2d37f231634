"""Oracle: LLaMA decoder stack with the reference's MoE wiring (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Restates transformers==4.31.0 ``models/llama/modeling_llama.py`` (LlamaRMSNorm, LlamaRotaryEmbedding,
apply_rotary_pos_emb, LlamaAttention eager path, LlamaMLP, _prepare_decoder_attention_mask) — un-vendored, pinned at
/root/reference/requirements.txt:137, numerics per SURVEY.md App. A.1 — wired the way the reference wires it:
  decoder layer  model/medplib/model/language_model/medplib_moe_llama.py:110-162
  model forward  model/medplib/model/language_model/medplib_moe_llama.py:165-305
  causal-LM tail model/medplib/model/language_model/medplib_moe_llama.py:381-421
Weights are looked up in a state dict by the reference's parameter names (prefix ``model.layers.{i}.``).
"""
import math

import torch
import torch.nn.functional as F

from . import moe as _moe


def rmsnorm(x, weight, eps):
    """LlamaRMSNorm.forward (4.31): fp32 variance, cast back to the input dtype, then multiply by weight."""
    in_dtype = x.dtype
    x32 = x.to(torch.float32)
    var = x32.pow(2).mean(-1, keepdim=True)
    x32 = x32 * torch.rsqrt(var + eps)
    return weight * x32.to(in_dtype)


def rope_tables(head_dim, max_pos, theta, dtype):
    """LlamaRotaryEmbedding._set_cos_sin_cache (4.31): fp32 angles, tables stored/cast to the run dtype."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(max_pos, dtype=torch.float32)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q, k, cos, sin, position_ids):
    """apply_rotary_pos_emb (4.31). q,k [B,H,T,d]; cos/sin [max_pos,d]; position_ids [B or 1, T]."""
    cos = cos[position_ids].unsqueeze(1)
    sin = sin[position_ids].unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def decoder_attention_mask(attention_mask, T, past, dtype):
    """LlamaModel._prepare_decoder_attention_mask (4.31): additive [B,1,T,past+T], finfo.min where masked."""
    B = attention_mask.shape[0]
    minv = torch.finfo(dtype).min
    total = past + T
    m = torch.zeros((T, total), dtype=dtype)
    if T > 1:
        causal = torch.full((T, T), minv, dtype=dtype)
        causal = causal.masked_fill(torch.arange(T)[None, :] <= torch.arange(T)[:, None], 0)
        m[:, past:] = causal
    m = m[None, None].expand(B, 1, T, total)
    pad = (1.0 - attention_mask[:, None, None, :].to(dtype)).expand(B, 1, T, total)
    pad = pad.masked_fill(pad.to(torch.bool), minv)
    # 4.31 adds the two additive masks (finfo.min + finfo.min overflows to -inf in bf16; softmax treats both alike
    # except for fully masked rows, which the reference never produces for real tokens)
    return (m + pad).clamp(min=minv)


def linear(sd, name, x):
    """nn.Linear (bias-free), or — when the state dict carries adapter weights for it — peft==0.10.0
    ``tuners/lora/layer.py::Linear.forward`` (un-vendored, pinned at /root/reference/requirements.txt:80; call site
    train_ds_medplib.py:294-302): result = base(x) + lora_B(lora_A(dropout(x))) * scaling. Dropout is applied only
    when the state dict injects a keep mask for this module (``<name>.lora_dropout_mask`` [rows, in_features], with
    ``lora_dropout_p``): x * mask / (1 - p), nn.Dropout's training-mode arithmetic with the random draw made explicit."""
    y = F.linear(x, sd[name + ".weight"])
    a = sd.get(name + ".lora_A.default.weight")
    if a is not None:
        xd = x
        mask = sd.get(name + ".lora_dropout_mask")
        if mask is not None:
            rows = x.numel() // x.shape[-1]
            xd = x * mask[:rows].reshape(x.shape).to(x.dtype) / (1.0 - sd["lora_dropout_p"])
        y = y + F.linear(F.linear(xd, a), sd[name + ".lora_B.default.weight"]) * sd["lora_scaling"]
    return y


def attention(sd, prefix, x, cfg, mask, position_ids, past_kv, cos, sin):
    """LlamaAttention.forward eager path (4.31). Returns (out, (k, v)) with k,v [B,H,T_total,d]."""
    B, T, D = x.shape
    H, hd = cfg["num_heads"], cfg["hidden_size"] // cfg["num_heads"]
    q = linear(sd, prefix + "self_attn.q_proj", x).view(B, T, H, hd).transpose(1, 2)
    k = linear(sd, prefix + "self_attn.k_proj", x).view(B, T, H, hd).transpose(1, 2)
    v = linear(sd, prefix + "self_attn.v_proj", x).view(B, T, H, hd).transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin, position_ids)
    if past_kv is not None:
        k = torch.cat([past_kv[0], k], dim=2)
        v = torch.cat([past_kv[1], v], dim=2)
    w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd)
    w = w + mask
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v).transpose(1, 2).reshape(B, T, D)
    return linear(sd, prefix + "self_attn.o_proj", o), (k, v)


def mlp(sd, p, x):
    """LlamaMLP.forward: down(silu(gate(x)) * up(x)); p = module path incl. trailing dot."""
    return linear(sd, p + "down_proj", F.silu(linear(sd, p + "gate_proj", x)) * linear(sd, p + "up_proj", x))


def layer_mlp(sd, prefix, x, cfg, training, rts_uniform=None, gumbel=None):
    """self.mlp of one decoder layer: DeepSpeed MoE when the layer has a gate, plain LlamaMLP otherwise.

    Returns (out, l_aux or None, exp_counts or None, router logits or None)."""
    wg_key = prefix + "mlp.deepspeed_moe.gate.wg.weight"
    if wg_key not in sd:
        return mlp(sd, prefix + "mlp.", x), None, None, None
    m = cfg["moe"]
    E = sd[wg_key].shape[0]
    experts = [f"{prefix}mlp.deepspeed_moe.experts.deepspeed_experts.{e}." for e in range(E)]
    cf = m["capacity_factor"] if training else m["eval_capacity_factor"]
    out, l_aux, exp_counts, logits = _moe.moe_layer(
        x, sd[wg_key], [lambda t, p=p: mlp(sd, p, t) for p in experts], k=m["top_k_experts"], capacity_factor=cf,
        min_capacity=m["min_capacity"], rts_uniform=rts_uniform, gumbel=gumbel)
    return out, l_aux, exp_counts, logits


def decoder_layer(sd, i, x, cfg, mask, position_ids, past_kv, cos, sin, training=False, rts_uniform=None):
    """MoELlamaDecoderLayer_forward (medplib_moe_llama.py:110-162)."""
    prefix = f"model.layers.{i}."
    eps = cfg["rms_norm_eps"]
    residual = x
    h = rmsnorm(x, sd[prefix + "input_layernorm.weight"], eps)
    h, kv = attention(sd, prefix, h, cfg, mask, position_ids, past_kv, cos, sin)
    x = residual + h
    residual = x
    h = rmsnorm(x, sd[prefix + "post_attention_layernorm.weight"], eps)
    # the injected per-layer noise is what the gating of this layer consumes: RTS uniforms (top-1) / Gumbel noise (top-2)
    top2 = (cfg.get("moe") or {}).get("top_k_experts", 1) == 2
    h, l_aux, exp_counts, logits = layer_mlp(sd, prefix, h, cfg, training, None if top2 else rts_uniform,
                                             rts_uniform if top2 else None)
    x = residual + h
    return x, kv, l_aux, exp_counts, logits


def model_forward(sd, cfg, inputs_embeds, attention_mask=None, past_key_values=None, training=False,
                  rts_uniforms=None):
    """MoELlamaModel_forward (medplib_moe_llama.py:165-305) on inputs_embeds [B,T,D].

    Returns dict(last_hidden_state, hidden_states (L+1 tuple), past_key_values, moe_losses, exp_counts, gate_logits).
    """
    B, T, D = inputs_embeds.shape
    L = cfg["num_layers"]
    hd = D // cfg["num_heads"]
    past = past_key_values[0][0].shape[2] if past_key_values is not None else 0
    position_ids = torch.arange(past, past + T).unsqueeze(0)
    if attention_mask is None:
        attention_mask = torch.ones((B, past + T), dtype=torch.bool)
    dtype = inputs_embeds.dtype
    mask = decoder_attention_mask(attention_mask, T, past, dtype)
    cos, sin = rope_tables(hd, max(cfg.get("max_position_embeddings", 4096), past + T), cfg.get("rope_theta", 1e4),
                           dtype)
    x = inputs_embeds
    hidden, kvs, losses, counts, glogits = [], [], [], [], []
    for i in range(L):
        hidden.append(x)
        pkv = past_key_values[i] if past_key_values is not None else None
        u = rts_uniforms[i] if rts_uniforms is not None else None
        x, kv, l_aux, ec, lg = decoder_layer(sd, i, x, cfg, mask, position_ids, pkv, cos, sin, training, u)
        kvs.append(kv)
        if l_aux is not None:
            losses.append(l_aux)
            counts.append(ec)
            glogits.append(lg)
    x = rmsnorm(x, sd["model.norm.weight"], cfg["rms_norm_eps"])
    hidden.append(x)
    return dict(last_hidden_state=x, hidden_states=tuple(hidden), past_key_values=kvs, moe_losses=losses,
                exp_counts=counts, gate_logits=glogits)


def causal_lm_tail(sd, cfg, hidden, labels=None, moe_losses=()):
    """MedPLIBMoELlamaForCausalLM.forward tail (medplib_moe_llama.py:381-421): fp32 logits, shifted CE over the
    samples that have at least one valid label, plus router_aux_loss_coef * sum(l_aux)."""
    logits = F.linear(hidden, sd["lm_head.weight"]).float()
    loss = None
    if labels is not None:
        shift_logits = logits[..., :-1, :].contiguous()
        shift_labels = labels[..., 1:].contiguous()
        keep = (shift_labels != -100).any(dim=1)
        shift_logits = shift_logits[keep].view(-1, logits.shape[-1])
        shift_labels = shift_labels[keep].view(-1)
        loss = F.cross_entropy(shift_logits, shift_labels)
    moe_loss = None
    if len(moe_losses) > 0:
        moe_loss = cfg["moe"]["router_aux_loss_coef"] * sum(moe_losses)
        if loss is not None:
            loss = loss + moe_loss
    return logits, loss, moe_loss


def greedy_generate(sd, cfg, inputs_embeds, attention_mask, embed_fn, max_new_tokens, eos_token_id=None,
                    forced_tokens=None):
    """transformers 4.31 ``generation/utils.py::greedy_search`` with use_cache=True and output_hidden_states=True,
    as driven by MedPLIB.evaluate (model/MedPLIB.py:592-610): prefill over the spliced embeddings, then one token
    per step through the KV cache (attention mask rebuilt as ones, medplib_arch.py:232-244).

    embed_fn(token_ids [B,1]) -> [B,1,D]. forced_tokens {step: id} overrides the argmax (bench config injects <SEG>).
    Returns (new_tokens [B, n], last_hidden [B, T+n-1, D] = concatenated last-layer hidden states, step logits).
    """
    out = model_forward(sd, cfg, inputs_embeds, attention_mask)
    hiddens = [out["last_hidden_state"]]
    kv = out["past_key_values"]
    logits = F.linear(out["last_hidden_state"][:, -1:], sd["lm_head.weight"]).float()
    new_tokens, step_logits = [], [logits[:, 0]]
    B = inputs_embeds.shape[0]
    for step in range(max_new_tokens):
        nxt = logits[:, -1].argmax(-1)
        if forced_tokens is not None and step in forced_tokens:
            nxt = torch.full_like(nxt, forced_tokens[step])
        new_tokens.append(nxt)
        if eos_token_id is not None and bool((nxt == eos_token_id).all()):
            break
        if step == max_new_tokens - 1:
            break
        x = embed_fn(nxt.view(B, 1))
        past = kv[0][0].shape[2]
        out = model_forward(sd, cfg, x, torch.ones((B, past + 1), dtype=torch.bool), kv)
        kv = out["past_key_values"]
        hiddens.append(out["last_hidden_state"])
        logits = F.linear(out["last_hidden_state"], sd["lm_head.weight"]).float()
        step_logits.append(logits[:, 0])
    return torch.stack(new_tokens, 1), torch.cat(hiddens, 1), step_logits
