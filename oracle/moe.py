"""Oracle: DeepSpeed-MoE layer (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).  PARITY UNPINNED against DeepSpeed.

Restates deepspeed==0.13.1 (un-vendored; pinned at /root/reference/requirements.txt:22; absent from this image, so it
cannot be executed to pin this file): ``deepspeed/moe/sharded_moe.py`` {_capacity, _top_idx, top1gating, top2gating,
TopKGate.forward, MOELayer.forward}, ``deepspeed/moe/experts.py::Experts.forward`` and ``deepspeed/moe/layer.py::MoE``
(use_residual=False, ep_size=1, noisy_gate_policy=None, drop_tokens=True, use_rts=True, use_tutel=False — the values
the reference passes at model/MedPLIB.py:253-263 and medplib_moe_llama.py:604-614,664-674). Numerics per SURVEY.md
App. A.3. Call site being reproduced: medplib_moe_llama.py:141-147 (``self.mlp(hidden_states)`` returning
``(out, l_aux, exp_counts)``).
"""
import math

import torch
import torch.nn.functional as F


def capacity(num_tokens, num_experts, capacity_factor, min_capacity):
    """_capacity: max(ceil(S / E * cf), min_capacity)."""
    c = int(math.ceil((num_tokens / num_experts) * capacity_factor))
    return max(c, int(min_capacity))


def _keep_mask(mask1, rand, C):
    """mask1 * zeros.scatter_(0, topk(mask1_rand, C, dim=0).indices, 1): keep at most C tokens per expert."""
    S = mask1.shape[0]
    k = min(C, S)
    top_idx = torch.topk(rand, k=k, dim=0).indices
    return mask1 * torch.zeros_like(mask1).scatter_(0, top_idx, 1)


_FORCED = None  # test hook, see forced_routing()


class forced_routing:
    """Test hook (like the injected RTS uniforms): inside this context the i-th top1gating call takes its expert index
    from decisions[i] ([S] int64) instead of the argmax. Used by the full-width parity tests to evaluate the oracle
    under the GPU path's routing decisions: a token whose two router logits are a near-tie flips under any change of
    rounding (the reference's own bf16 and fp32 runs disagree on such tokens), which says nothing about the arithmetic
    being compared; the tests separately require the decisions to agree wherever the margin exceeds the noise."""

    def __init__(self, decisions):
        self.decisions = list(decisions)

    def __enter__(self):
        global _FORCED
        _FORCED = iter(self.decisions)
        return self

    def __exit__(self, *a):
        global _FORCED
        _FORCED = None


def top1gating(logits, capacity_factor, min_capacity, rts_uniform=None):
    """top1gating. logits fp32 [S,E]. rts_uniform: the U(0,1) sample DeepSpeed draws for Random Token Selection
    ([S,E]); None = no overflow expected (position order is used, which is what topk does when nothing overflows).

    Returns l_aux, gate value per token [S] (0 if dropped), expert index [S], slot [S] (-1 if dropped), C, exp_counts.
    """
    S, E = logits.shape
    gates = F.softmax(logits, dim=1)
    C = capacity(S, E, capacity_factor, min_capacity)
    idx = torch.argmax(gates, dim=1)
    if _FORCED is not None:
        idx = next(_FORCED).to(idx.device).long().reshape(idx.shape)
    mask1 = F.one_hot(idx, num_classes=E)
    exp_counts = mask1.sum(0)
    me = gates.mean(0)
    ce = mask1.float().mean(0)
    l_aux = (me * ce).sum() * E
    if rts_uniform is not None:
        rand = mask1 * rts_uniform
    else:
        # deterministic stand-in: earlier tokens win (exact when no expert overflows)
        rand = mask1.float() * (2.0 - torch.arange(S, dtype=torch.float32)[:, None] / max(S, 1))
    assert S >= min_capacity
    mask1 = _keep_mask(mask1, rand, C)
    loc = torch.cumsum(mask1, dim=0) - 1
    loc_s = (loc * mask1).sum(1)
    kept = mask1.sum(1) > 0
    gate_s = (gates * mask1.float()).sum(1)
    slot = torch.where(kept, loc_s, torch.full_like(loc_s, -1))
    return l_aux, gate_s, idx, slot, C, exp_counts


def top2gating(logits, capacity_factor, min_capacity, gumbel=None):
    """top2gating (0.13.1): second expert = argmax of (logits + Gumbel noise) with the first masked out; capacity
    2*cf*S/E; tokens past capacity dropped in position order; the two gate values renormalised by their sum."""
    S, E = logits.shape
    gates = F.softmax(logits, dim=1)
    C = capacity(S, E, capacity_factor * 2, min_capacity)
    idx1 = torch.argmax(gates, dim=1)
    mask1 = F.one_hot(idx1, num_classes=E)
    noisy = logits + (gumbel if gumbel is not None else 0.0)
    noisy = noisy.masked_fill(mask1.bool(), float("-inf"))
    idx2 = torch.argmax(noisy, dim=1)
    mask2 = F.one_hot(idx2, num_classes=E)
    loc1 = torch.cumsum(mask1, dim=0) - 1
    loc2 = torch.cumsum(mask2, dim=0) - 1 + mask1.sum(0, keepdim=True)
    exp_counts = mask1.sum(0)
    l_aux = (gates.mean(0) * mask1.float().mean(0)).mean() * E * E
    mask1 = mask1 * (loc1 < C)
    mask2 = mask2 * (loc2 < C)
    loc1_s = (loc1 * mask1).sum(1)
    loc2_s = (loc2 * mask2).sum(1)
    g1 = (gates * mask1.float()).sum(1)
    g2 = (gates * mask2.float()).sum(1)
    denom = torch.clamp(g1 + g2, min=torch.finfo(gates.dtype).eps)
    g1, g2 = g1 / denom, g2 / denom
    slot1 = torch.where(mask1.sum(1) > 0, loc1_s, torch.full_like(loc1_s, -1))
    slot2 = torch.where(mask2.sum(1) > 0, loc2_s, torch.full_like(loc2_s, -1))
    return l_aux, (g1, g2), (idx1, idx2), (slot1, slot2), C, exp_counts


def moe_layer(x, wg, experts, k, capacity_factor, min_capacity, rts_uniform=None, gumbel=None):
    """MoE.forward -> MOELayer.forward with ep_size=1. x [B,T,D] (any float dtype); wg fp32 [E,D] (TopKGate keeps the
    gate in fp32 and casts its input up); experts: list of callables on [n,D].

    DeepSpeed materialises dispatch/combine as dense one-hot einsums 'sec,sm->ecm' / 'sec,ecm->sm'; with a one-hot
    dispatch mask those are exactly a row gather into [E,C,D] (zeros elsewhere) and, on the way back,
    out[s] = sum_j type_as(gate_j[s], x) * expert_out[e_j, slot_j] accumulated in the input dtype's matmul precision.
    Returns (out [B,T,D], l_aux, exp_counts, logits).
    """
    shape = x.shape
    D = shape[-1]
    xs = x.reshape(-1, D)
    S = xs.shape[0]
    logits = F.linear(xs.float(), wg.float())
    E = wg.shape[0]
    if k == 1:
        l_aux, g, idx, slot, C, exp_counts = top1gating(logits, capacity_factor, min_capacity, rts_uniform)
        routes = [(g, idx, slot)]
    else:
        l_aux, gs, idxs, slots, C, exp_counts = top2gating(logits, capacity_factor, min_capacity, gumbel)
        routes = list(zip(gs, idxs, slots))
    dispatched = torch.zeros((E, C, D), dtype=x.dtype)
    for g, idx, slot in routes:
        keep = slot >= 0
        dispatched[idx[keep], slot[keep]] = xs[keep]
    expert_out = torch.stack([experts[e](dispatched[e]) for e in range(E)], 0)
    out32 = torch.zeros((S, D), dtype=torch.float32)
    for g, idx, slot in routes:
        keep = slot >= 0
        w = g.to(x.dtype)  # combine_weights.type_as(input)
        out32[keep] += w[keep].float()[:, None] * expert_out[idx[keep], slot[keep]].float()
    return out32.to(x.dtype).reshape(shape), l_aux, exp_counts, logits
