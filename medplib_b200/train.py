"""Train step of the hot path: LoRA adapters, the flat fp32 gradient arena, the LLaMA-MoE
stack with a hand-scheduled backward, the lm_head + shifted cross-entropy tail, bucketed data-parallel all-reduce and
fused AdamW. Replaces what peft 0.10 + torch autograd + DeepSpeed ZeRO-2 run for train_ds_medplib.py:294-302,398-439,
599-625 around model/medplib/model/language_model/medplib_moe_llama.py:110-438.

Design (B200-first, 180 GB per GPU):
* No gradient checkpointing: the reference recomputes every decoder layer (medplib_moe_llama.py:252-263) because 80 GB
  cards force it; here the activations of all 32 layers (~1 GB per layer at 8x639 tokens) stay resident, so a step is
  fwd + bwd, not fwd + recompute + bwd.
* Frozen base weights are kept twice (W and W^T, made once with mpl_transpose_bf16) so every dX = dY.W contraction is
  the same K-major tcgen05 GEMM as the forward; q,k,v share one [D,3D] transposed matrix (one dgrad GEMM with K = 3D).
* All trainable parameters' gradients live in ONE flat fp32 arena in backward-completion order; kernels accumulate
  into it directly (no per-parameter .grad tensors, no bf16 gradient rounding), the data-parallel all-reduce runs over
  arena buckets on NCCL as the backward passes them, and AdamW + gradient clipping run over the arena.
* torch.autograd is only the tape between three coarse nodes (stack, lm_head+CE, mask head); every arithmetic kernel
  is ours. There is no eager fallback: CPU tensors raise MplError.
"""
import math
import os

import torch
import torch.nn as nn

from . import _lib, ops
from . import train_ops as T
from .model import modules as M

bf16 = torch.bfloat16
f32 = torch.float32
# LoRA up-projections in the epilogue of the base GEMM (mpl_gemm_args.lora_*) instead of a separate pass over its output
LORA_FUSE = os.environ.get("MPL_LORA_FUSE", "1") != "0"
# ... and, where every adapter of an output qualifies (rank 8, no lora_dropout), inside the GEMM's accumulator: one extra
# k-block [u_0 | u_1 | ..] x [s B_0 | s B_1 | ..]^T on the tensor cores (mpl_gemm_args.ext_a / ext_b)
LORA_EXT = os.environ.get("MPL_LORA_EXT", "1") != "0"
SILU_BWD_FUSE = os.environ.get("MPL_SILU_BWD_FUSE", "0") == "1"
LORA_EXCLUDE = ("visual_model", "vision_tower", "mm_projector")  # train_ds_medplib.py:272-281


# ---------------------------------------------------------------------------------------------------- LoRA adapters
def find_linear_layers(model, lora_target_modules):
    """train_ds_medplib.py:265-291: names of nn.Linear modules matched by substring, vision parts excluded."""
    names = set()
    for name, module in model.named_modules():
        if isinstance(module, nn.Linear) and all(x not in name for x in LORA_EXCLUDE) \
                and any(x in name for x in lora_target_modules) and "lora_" not in name:
            names.add(name)
    return sorted(names)


def attach_lora(model, r=8, lora_alpha=16, lora_dropout=0.0, target_modules=("q_proj", "v_proj")):
    """What get_peft_model(model, LoraConfig(r, lora_alpha, target_modules, lora_dropout, bias="none")) does to the
    module tree (train_ds_medplib.py:294-302), without peft: every matched nn.Linear gets ``lora_A.default`` /
    ``lora_B.default`` Linear children (peft's names; A kaiming-uniform, B zero), every other parameter is frozen.
    The matched Linear keeps its own ``weight`` (peft moves it to ``base_layer.weight`` — see INTEGRATION.md)."""
    if not 0.0 <= float(lora_dropout) < 1.0:
        raise _lib.MplError("lora_dropout must be in [0, 1)")
    if isinstance(target_modules, str):
        target_modules = target_modules.split(",")
    names = find_linear_layers(model, list(target_modules))
    for p in model.parameters():
        p.requires_grad = False
    mods = dict(model.named_modules())
    for name in names:
        lin = mods[name]
        w = lin.weight
        A = nn.Linear(lin.in_features, r, bias=False, device=w.device, dtype=w.dtype)
        Bm = nn.Linear(r, lin.out_features, bias=False, device=w.device, dtype=w.dtype)
        nn.init.kaiming_uniform_(A.weight, a=math.sqrt(5))
        nn.init.zeros_(Bm.weight)
        lin.lora_A = nn.ModuleDict({"default": A})
        lin.lora_B = nn.ModuleDict({"default": Bm})
        lin.scaling = {"default": lora_alpha / r}
        lin.r = {"default": r}
        lin.lora_dropout_p = float(lora_dropout)
        lin.lora_name = name
    if hasattr(model, "refresh_engines"):
        model.refresh_engines()
    return names


def set_trainable(model, sft_modules):
    """train_ds_medplib.py:316-326: parameters whose name contains one of the substrings become trainable."""
    if isinstance(sft_modules, str):
        sft_modules = [s for s in sft_modules.split(",") if s]
    for n, p in model.named_parameters():
        if any(x in n for x in sft_modules):
            p.requires_grad = True


class _Lora:
    __slots__ = ("A", "B", "s", "gA", "gB", "p", "name", "BT")

    def __init__(self, lin, arena):
        self.A, self.B = lin.lora_A["default"].weight, lin.lora_B["default"].weight
        self.s = float(lin.scaling["default"])
        p = getattr(lin, "lora_dropout_p", None)
        if p is None:  # a real peft lora.Linear: lora_dropout is a ModuleDict of nn.Dropout / nn.Identity
            drop = getattr(lin, "lora_dropout", None)
            p = getattr(drop["default"], "p", 0.0) if isinstance(drop, nn.ModuleDict) and "default" in drop else 0.0
        self.p = float(p)
        self.name = getattr(lin, "lora_name", "")
        self.gA, self.gB = arena.of(self.A), arena.of(self.B)
        if self.A.shape[0] > 16:
            raise _lib.MplError("LoRA rank > 16 is not built (the reference trains with r = 8)")


def _lora_of(lin, arena):
    return _Lora(lin, arena) if hasattr(lin, "lora_A") else None


# ------------------------------------------------------------------------------------------------- gradient arena
class GradArena:
    """One flat fp32 buffer holding the gradient of every trainable parameter, in the order given (the order the
    backward completes them, so bucket i can be all-reduced while bucket i+1 is still being produced)."""

    ALIGN = 64

    def __init__(self, named_params, device):
        self.names, self.params, self.offsets = [], [], []
        off = 0
        for n, p in named_params:
            self.names.append(n)
            self.params.append(p)
            self.offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = off
        self.flat = torch.zeros(max(off, 1), dtype=f32, device=device)
        self._views = {id(p): self.flat[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)}
        self._end = {id(p): o + p.numel() for p, o in zip(self.params, self.offsets)}

    def of(self, p):
        """fp32 gradient view of parameter p, or None when p is frozen."""
        return self._views.get(id(p)) if p is not None else None

    def end_of(self, p):
        return self._end[id(p)]

    def zero_(self):
        self.flat.zero_()

    def grads(self):
        return {n: self._views[id(p)] for n, p in zip(self.names, self.params)}

    def export_grads(self):
        """Materialise p.grad (param dtype) for code that expects torch-style gradients (DeepSpeed / torch.optim)."""
        for p in self.params:
            p.grad = self._views[id(p)].to(p.dtype)


class BucketReducer:
    """Data-parallel gradient exchange: all-reduce(mean) of the arena in fixed-size buckets, each
    launched asynchronously as soon as the backward has passed its end, on the process group's own stream."""

    def __init__(self, arena, bucket_elems=64 << 20, group=None, wire_dtype=None):
        """wire_dtype: the dtype on the wire. Default bf16 on CUDA -- what the reference exchanges (DeepSpeed reduces the
        bf16 gradients of a bf16 engine, train_ds_medplib.py:408-419): half the bytes of the fp32 arena, 636 MB per step
        for the stage-4 trainable set as SURVEY 2.2 counts it; the sum lands back in the fp32 arena. fp32 elsewhere."""
        import torch.distributed as dist
        self.dist, self.arena, self.group = dist, arena, group
        if wire_dtype is None:
            wire_dtype = bf16 if arena.flat.is_cuda and os.environ.get("MPL_DP_WIRE", "bf16") == "bf16" else f32
        self.wire = wire_dtype
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.on else 1
        n = arena.numel
        self.bounds = [(a, min(a + bucket_elems, n)) for a in range(0, n, bucket_elems)] or [(0, 0)]
        self.owner = None  # Trainer (for the gradient-accumulation guard)
        self.reset()

    def reset(self):
        self.next, self.work = 0, []

    def ready(self, upto):
        """Everything in arena[:upto] is final: launch the buckets that end at or before `upto`."""
        if not self.on:
            return
        while self.next < len(self.bounds) and self.bounds[self.next][1] <= upto:
            if self.next == 0 and self.owner is not None:
                self.launch_micro_step = self.owner.micro_steps
            a, b = self.bounds[self.next]
            buf = self.arena.flat[a:b] if self.wire == f32 else self.arena.flat[a:b].to(self.wire)
            self.work.append((self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True),
                              a, b, buf))
            self.next += 1

    def finish(self):
        """Launch what is left and make the current stream wait for every bucket. Returns 1/world (the mean factor the
        optimizer folds into its gradient scale)."""
        self.ready(self.arena.numel)
        for w, a, b, buf in self.work:
            w.wait()
            if self.wire != f32:
                self.arena.flat[a:b].copy_(buf)
        self.reset()
        return 1.0 / self.world


class FusedAdamW:
    """AdamW over the arena (train_ds_medplib.py:398-411: betas (0.9, 0.95), weight decay 0, clipping 1.0) with fp32
    master weights + moments, writing the bf16 / fp32 parameters in place. Clipping uses the global gradient norm
    computed on the device (no .item())."""

    def __init__(self, arena, lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0):
        self.arena, self.lr, self.betas, self.eps, self.wd, self.max_norm = arena, lr, betas, eps, weight_decay, \
            max_grad_norm
        dev = arena.flat.device
        self.master = torch.zeros_like(arena.flat)
        for p, o in zip(arena.params, arena.offsets):
            self.master[o:o + p.numel()].copy_(p.detach().reshape(-1).float())
        self.m = torch.zeros_like(arena.flat)
        self.v = torch.zeros_like(arena.flat)
        self.sumsq = torch.zeros(1, dtype=f32, device=dev)
        self.t = 0
        # chunk table of the one-launch update: every 4096-element chunk of the arena -> its parameter
        rows = []
        for p, o in zip(arena.params, arena.offsets):
            if not p.is_contiguous():
                raise _lib.MplError("trainable parameters must be contiguous")
            n, bf = p.numel(), int(p.dtype == bf16)
            for c in range(0, n, 4096):
                rows.append((o + c, min(4096, n - c), p.data_ptr(), bf | (c << 1)))
        self.chunks = torch.tensor(rows, dtype=torch.int64).to(dev)
        self._ptrs = [p.data_ptr() for p in arena.params]

    def step(self, grad_scale=1.0, lr=None):
        a = self.arena
        self.t += 1
        sumsq = None
        if self.max_norm and self.max_norm > 0:
            self.sumsq.zero_()
            T.sumsq(a.flat, self.sumsq)
            sumsq = self.sumsq
        lr = self.lr if lr is None else lr
        if [p.data_ptr() for p in a.params] != self._ptrs:
            raise _lib.MplError("a trainable parameter's storage moved since the optimizer was built; call "
                                "model.trainer(...) again")
        T.adamw_multi(self.master, self.m, self.v, a.flat, self.chunks, lr, self.betas[0], self.betas[1], self.eps,
                      self.wd, self.t, sumsq_dev=sumsq, max_norm=self.max_norm or 0.0, grad_scale=grad_scale)

    def grad_norm(self, grad_scale=1.0):
        return self.sumsq.sqrt() * grad_scale


# --------------------------------------------------------------------------------------- LLaMA-MoE stack, fwd + bwd
class _Layer:
    pass


def _frozen(w, what):
    if w.requires_grad:
        raise _lib.MplError(f"full fine-tuning of {what} is not built: the reference trains the decoder through LoRA "
                            "(train_ds_medplib.py:294-302); keep the base weight frozen")
    return w


class _ExtSite:
    """The adapters feeding ONE base GEMM as its extension k-block: weight-side operand `b` (bf16 [n_mats * N, 64], adapter t
    in columns [8t, 8t + 8) of the rows of the matrix it feeds, zero elsewhere) + the pack records that fill it."""

    def __init__(self, n_mats, N, device):
        self.N = N
        self.b = torch.zeros((n_mats * N, 64), dtype=bf16, device=device)
        self.col = {}     # id(_Lora) -> first column
        self.items = []   # (src tensor, dst row offset, sn, sr, N, r, col, scale)

    def add(self, lo, mat, transposed):
        """transposed=False: forward, rows = s * lora_B [N, r]; True: dgrad, rows = lora_A^T (lora_A is [r, N])."""
        col = 8 * len(self.col)
        self.col[id(lo)] = col
        src = lo.A if transposed else lo.B
        r = lo.A.shape[0]
        if transposed:
            self.items.append((src, mat * self.N, src.stride(1), src.stride(0), self.N, r, col, 1.0))
        else:
            self.items.append((src, mat * self.N, src.stride(0), src.stride(1), self.N, r, col, lo.s))


class LlamaTrainStack:
    """Training forward (activations kept) and backward of the decoder stack of a MedPLIBForCausalLM."""

    def __init__(self, model, arena, reducer=None):
        cfg = model.config
        self.arena, self.reducer = arena, reducer
        self.D, self.H, self.F = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        self.hd = self.D // self.H
        self.eps = cfg.rms_norm_eps
        moe = getattr(cfg, "moe", None) or {}
        self.cf = float(moe.get("capacity_factor", 1.0) or 1.0)
        self.min_cap = int(moe.get("min_capacity", 0) or 0)
        self.top_k = int(moe.get("top_k_experts", 1) or 1)
        self.aux_coef = float(getattr(model, "router_aux_loss_coef", 0.0) or 0.0)
        self.use_rts = True  # DeepSpeed top1gating default (use_rts=True): random token selection on overflow
        self.debug = None  # dict: stage name -> tensor copies (tests/dev/debug_train.py)
        self.training = True          # lora_dropout is active (model.train()); Trainer keeps it in sync with the model
        self.dropout_masks = None     # tests: {adapted module name: keep mask [rows, in_features]} instead of torch.rand
        from .engine import rope_tables
        from .model.config import llama_dims
        dims = llama_dims(cfg)
        self.rope_len = int(dims.get("max_position_embeddings", 4096))
        self.cos, self.sin = rope_tables(self.hd, self.rope_len, dims.get("rope_theta", 1e4),
                                         model.lm_head.weight.device)
        self.norm_w = model.model.norm.weight
        self.layers = []
        for li, layer in enumerate(model.model.layers):
            L = _Layer()
            at = layer.self_attn
            L.ln1, L.ln2 = layer.input_layernorm.weight, layer.post_attention_layernorm.weight
            L.wq, L.wk, L.wv, L.wo = (_frozen(at.q_proj.weight, "q_proj"), _frozen(at.k_proj.weight, "k_proj"),
                                      _frozen(at.v_proj.weight, "v_proj"), _frozen(at.o_proj.weight, "o_proj"))
            L.lo = {n: _lora_of(getattr(at, n), arena) for n in ("q_proj", "k_proj", "v_proj", "o_proj")}
            D = self.D
            L.wqkvT = torch.empty((D, 3 * D), dtype=bf16, device=L.wq.device)
            for j, w in enumerate((L.wq, L.wk, L.wv)):
                T.transpose(w.detach(), out=L.wqkvT[:, j * D:(j + 1) * D])
            L.woT = T.transpose(L.wo.detach())
            if isinstance(layer.mlp, M.MoE):
                if self.top_k not in (1, 2):
                    raise _lib.MplError("DeepSpeed's TopKGate supports top-1 and top-2 gating only")
                L.wg = layer.mlp.deepspeed_moe.gate.wg.weight
                mlps = list(layer.mlp.deepspeed_moe.experts.deepspeed_experts)
            else:
                L.wg = None
                mlps = [layer.mlp]
            L.E = len(mlps)
            L.w_gate = [_frozen(m.gate_proj.weight, "gate_proj") for m in mlps]
            L.w_up = [_frozen(m.up_proj.weight, "up_proj") for m in mlps]
            L.w_down = [_frozen(m.down_proj.weight, "down_proj") for m in mlps]
            L.w_gateT = [T.transpose(w.detach()) for w in L.w_gate]
            L.w_upT = [T.transpose(w.detach()) for w in L.w_up]
            L.w_downT = [T.transpose(w.detach()) for w in L.w_down]
            L.lo_mlp = [{n: _lora_of(getattr(m, n), arena) for n in ("gate_proj", "up_proj", "down_proj")}
                        for m in mlps]
            # arena offset below which everything is final once this layer's backward is done
            ends = [arena.end_of(p) for p in layer.parameters() if arena.of(p) is not None]
            L.arena_end = max(ends) if ends else None
            self._ext_sites(L, D, self.F, L.wq.device)
            self.layers.append(L)
        self._ext_table()

    # ------------------------------------------------------------------ adapters as extension k-blocks
    @staticmethod
    def _ext_ok(los):
        return LORA_EXT and LORA_FUSE and len(los) > 0 and len(los) <= 8 and all(lo.A.shape[0] == 8 for lo in los)

    def _ext_sites(self, L, D, F_, dev):
        """Forward and dgrad sites of one layer (None where no adapter feeds the GEMM or one of them does not qualify)."""
        L.ext_f, L.ext_b = {}, {}

        def make(key, n_mats, N, pairs, transposed, table):
            los = [lo for _, lo in pairs]
            if not self._ext_ok(los):
                table[key] = None
                return
            site = _ExtSite(n_mats, N, dev)
            for mat, lo in pairs:
                site.add(lo, mat, transposed)
            table[key] = site

        qkv = [(j, L.lo[nm]) for j, nm in enumerate(("q_proj", "k_proj", "v_proj")) if L.lo[nm] is not None]
        make("qkv", 3, D, qkv, False, L.ext_f)
        make("qkv", 1, D, [(0, lo) for _, lo in qkv], True, L.ext_b)
        o = [(0, L.lo["o_proj"])] if L.lo["o_proj"] is not None else []
        make("o", 1, D, o, False, L.ext_f)
        make("o", 1, D, o, True, L.ext_b)
        for e in range(L.E):
            # gate and up as ONE dual GEMM (SiLU(gate) * up in its epilogue): both adapters in the same extension block
            gu = [(j, L.lo_mlp[e][nm]) for j, nm in enumerate(("gate_proj", "up_proj")) if L.lo_mlp[e][nm] is not None]
            make(("gateup", e), 2, F_, gu, False, L.ext_f)
            for nm, n_out, n_in in (("gate_proj", F_, D), ("up_proj", F_, D), ("down_proj", D, F_)):
                lo = L.lo_mlp[e][nm]
                one = [(0, lo)] if lo is not None else []
                make((nm, e), 1, n_out, one, False, L.ext_f)
                make((nm, e), 1, n_in, one, True, L.ext_b)

    def _ext_table(self):
        """Device table of pack records (48 bytes each, mpl_lora_pack) over every site of the model + its dirty flag."""
        import struct
        recs, keep = [], []
        for L in self.layers:
            for table in (L.ext_f, L.ext_b):
                for site in table.values():
                    if site is None:
                        continue
                    for src, row0, sn, sr, N, r, col, scale in site.items:
                        recs.append(struct.pack("<QQqqiiifqq", src.data_ptr(), site.b.data_ptr() + row0 * 64 * 2, sn, sr, N, r, col,
                                                scale, 64, 1))
                        keep.append(src)
            # lora_B^T [r, N] of every adapter (the down-projection du = s dY B of its backward), refreshed by the same launch
            los = [lo for lo in L.lo.values() if lo is not None] + [lo for d in L.lo_mlp for lo in d.values() if lo is not None]
            for lo in los:
                N, r = lo.B.shape
                lo.BT = torch.empty((r, N), dtype=bf16, device=lo.B.device)
                recs.append(struct.pack("<QQqqiiifqq", lo.B.data_ptr(), lo.BT.data_ptr(), lo.B.stride(0), lo.B.stride(1), N, r, 0,
                                        1.0, 1, N))
                keep.append(lo.BT)
        self._ext_n = len(recs)
        self._ext_keep = keep
        self._ext_items = None
        if recs:
            dev = self.layers[0].wq.device
            self._ext_items = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(dev)
        self.ext_dirty = True
        self._pads = {}

    def _ext_refresh(self):
        """(Re)build every weight-side extension operand from the current adapter weights: one launch, after each
        optimizer step (every forward when a foreign optimizer owns the weights)."""
        if self._ext_items is not None and self.ext_dirty:
            T.lora_pack(self._ext_items, self._ext_n)
            self.ext_dirty = False

    def _pad(self, rows, dev):
        """Activation-side extension operand bf16 [rows, 64]: zero once, columns [8t, 8t + 8) rewritten per use."""
        buf = self._pads.get(rows)
        if buf is None:
            buf = self._pads[rows] = torch.zeros((rows, 64), dtype=bf16, device=dev)
        return buf

    def _site_usable(self, site, los):
        return site is not None and not (self.training and any(lo.p > 0.0 for lo in los))

    def _linear_lora(self, x, weights, los, site, out=None, residual=None, m_dev=None, force=None):
        """y_i = x W_i^T (+ residual) with the adapters `los` = [(matrix index, _Lora)] of its outputs. Returns
        (result of ops.linear, {matrix index: saved for the backward})."""
        saved = {}
        multi = isinstance(weights, (list, tuple))
        if not los:
            return ops.linear(x, weights, out=out, residual=residual, m_dev=m_dev, force=force), saved
        if self._site_usable(site, [lo for _, lo in los]):
            buf = self._pad(x.shape[0], x.device)
            for mat, lo in los:
                saved[mat] = self._lora_a(lo, x, pad=(buf, site.col[id(lo)]))
            return ops.linear(x, weights, out=out, residual=residual, m_dev=m_dev, force="tc", ext=(buf, site.b)), saved
        terms, later = [], []
        for mat, lo in los:
            saved[mat] = self._lora_a(lo, x)
            (terms if len(terms) < 2 else later).append((mat, lo))
        y = ops.linear(x, weights, out=out, residual=residual if not later else None, m_dev=m_dev, force="tc",
                       lora=[self._lora_term(lo, saved[mat], mat) for mat, lo in terms])
        ys = y if multi else [y]
        for mat, lo in later:
            T.lora_up_add(ys[mat], saved[mat][0], lo.B, lo.s)
        if later and residual is not None:
            assert not multi
            y = ops.add(residual, y)
        return y, saved

    def _dgrad_lora(self, dy, wT, adapters, site, out=None, residual=None, m_dev=None, force=None, silu_bwd=None):
        """dx = dy W (+ residual) + sum over `adapters` = [(_Lora, x, saved, dy slice of that adapter)] of du A; their dA, dB
        go into the arena."""
        if not adapters:
            return ops.linear(dy, wT, out=out, residual=residual, m_dev=m_dev, force="tc" if silu_bwd is not None else force,
                              silu_bwd=silu_bwd)
        if self._site_usable(site, [a[0] for a in adapters]):
            buf = self._pad(dy.shape[0], dy.device)
            for lo, x, saved, dys in adapters:
                self._lora_bwd_pre(lo, x, saved, dys, pad=(buf, site.col[id(lo)]))
            return ops.linear(dy, wT, out=out, residual=residual, m_dev=m_dev, force="tc", ext=(buf, site.b),
                              silu_bwd=silu_bwd)
        assert silu_bwd is None
        terms, pend = [], []
        for lo, x, saved, dys in adapters:
            term, du = self._lora_bwd_pre(lo, x, saved, dys)
            if term is not None and len(terms) < 2:
                terms.append(term)
            else:
                pend.append((lo, saved, du))
        dx = ops.linear(dy, wT, out=out, residual=residual, m_dev=m_dev, force="tc" if terms else force, lora=terms or None)
        for lo, saved, du in pend:
            self._lora_bwd_dx(lo, saved, du, dx)
        return dx

    # ------------------------------------------------------------------ helpers
    def _lora_a(self, lo, x, pad=None):
        """a = dropout(x) A^T (bf16 [M, r]) in peft's rounding; returns what the backward needs: (a, dropped input or
        None, keep mask or None)."""
        xd = mask = None
        if lo.p > 0.0 and self.training:
            # peft: lora_B(lora_A(lora_dropout(x))) — every adapted Linear owns its dropout, masks are per call
            inj = self.dropout_masks.get(lo.name) if self.dropout_masks is not None else None
            if inj is not None:
                mask = inj[:x.shape[0]].to(torch.uint8).contiguous()
            else:
                mask = (torch.rand(x.shape, device=x.device) >= lo.p).to(torch.uint8)
            xd = T.mask_scale(x if x.is_contiguous() else x.contiguous(), mask, 1.0 / (1.0 - lo.p))
        a = T.lora_down(xd if xd is not None else x, lo.A, pad=pad)
        return (a, xd, mask)

    def _lora_fwd(self, lo, x, y):
        """y += s * (dropout(x) A^T) B^T as its own pass over y (the GEMM that produced y could not take the adapter in its
        epilogue)."""
        saved = self._lora_a(lo, x)
        T.lora_up_add(y, saved[0], lo.B, lo.s)
        return saved

    @staticmethod
    def _lora_term(lo, saved, mat=0):
        """The adapter as a fused epilogue term of the base GEMM: (u, b [N, r], scale, output matrix)."""
        return (saved[0], lo.B.detach(), lo.s, mat)

    def _lora_bwd_pre(self, lo, x, saved, dy, pad=None):
        """Everything of the adapter's backward except dx: dB, dA into the arena. Returns the epilogue term of the base
        dgrad GEMM, (du f32 [M, r], A^T [in, r], 1.0, 0), or None when lora_dropout is active (dx then needs the mask:
        _lora_bwd_dx)."""
        a, xd, mask = saved
        if lo.gB is not None:
            T.rank_wgrad(dy, a, lo.gB, scale=lo.s)
        bt = getattr(lo, "BT", None)  # (kept current by the per-step lora_pack launch)
        du = T.lora_down(dy, bt if bt is not None else T.transpose(lo.B.detach()), scale=lo.s, out_f32=True, pad=pad)
        if lo.gA is not None:
            T.rank_wgrad(xd if xd is not None else x, du, lo.gA, transposed=True)
        if mask is not None or pad is not None:
            return None, du
        return (du, T.transpose(lo.A.detach()), 1.0, 0), du

    def _lora_bwd_dx(self, lo, saved, du, dx):
        """dx += dropout'(du A) as its own pass (lora_dropout active, or no base GEMM to fuse into)."""
        mask = saved[2]
        if mask is None:
            T.lora_up_add(dx, du, lo.A, 1.0, transposed=True)
        else:
            tmp = torch.zeros((dx.shape[0], dx.shape[1]), dtype=bf16, device=dx.device)
            T.lora_up_add(tmp, du, lo.A, 1.0, transposed=True)
            T.mask_scale(tmp, mask, 1.0 / (1.0 - lo.p), out=dx, accumulate=True)

    def _lora_bwd(self, lo, x, saved, dy, dx):
        """Gradients of y = ... + s (dropout(x) A^T) B^T: dB, dA into the arena, dx += dropout'(du A)."""
        _, du = self._lora_bwd_pre(lo, x, saved, dy)
        self._lora_bwd_dx(lo, saved, du, dx)

    def capacity(self, S, E):
        return ops.moe_capacity(S, E, self.cf, self.min_cap, self.top_k)

    # ------------------------------------------------------------------ forward
    def forward(self, x, B, Tn, kv_mask=None, moe_noise=None):
        """x bf16 [B*Tn, D] spliced input embeddings. Returns (hidden [S,D] after the final norm, l_aux [n_moe] f32,
        gate_logits list, saved)."""
        D, H, hd, F_, eps = self.D, self.H, self.hd, self.F, self.eps
        S = B * Tn
        dev = x.device
        if Tn > self.rope_len:
            raise _lib.MplError("sequence longer than max_position_embeddings")
        cos, sin = self.cos, self.sin
        scale = 1.0 / math.sqrt(hd)
        km = kv_mask.to(torch.uint8).contiguous() if kv_mask is not None else None
        saved, l_aux, gate_logits = [], [], []
        self.ext_dirty = True  # (one ~50 us launch per step: cheaper than tracking who may have touched the adapters)
        self._ext_refresh()
        for li, L in enumerate(self.layers):
            sv = {"x": x}
            n1 = ops.rmsnorm(x, L.ln1, eps)
            qkv = torch.empty((S, 3 * D), dtype=bf16, device=dev)
            views = [qkv[:, j * D:(j + 1) * D] for j in range(3)]
            # adapters ride inside the base GEMM: as an extension k-block where they qualify, else in its epilogue
            los = [(j, L.lo[nm]) for j, nm in enumerate(("q_proj", "k_proj", "v_proj")) if L.lo[nm] is not None]
            _, sav = self._linear_lora(n1, [L.wq, L.wk, L.wv], los, L.ext_f["qkv"], out=views)
            for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                if j in sav:
                    sv["a_" + nm] = sav[j]
            q5 = qkv.view(B, Tn, 3, H, hd)
            q, k, v = q5[:, :, 0], q5[:, :, 1], q5[:, :, 2]
            ops.rope_kv(q, k, None, cos, sin, 0)
            o, lse = T.attention_fwd_lse(q, k, v, scale, causal=True, kv_mask=km)
            o2 = o.view(S, D)
            h1, sav = self._linear_lora(o2, L.wo, [(0, L.lo["o_proj"])] if L.lo["o_proj"] is not None else [],
                                        L.ext_f["o"], residual=x)
            if 0 in sav:
                sv["a_o_proj"] = sav[0]
            n2 = ops.rmsnorm(h1, L.ln2, eps)
            sv.update(n1=n1, qkv=qkv, o=o, lse=lse, h1=h1, n2=n2)
            E = L.E
            if L.wg is not None:
                C = self.capacity(S, E)
                noise = moe_noise[li] if moe_noise is not None else None
                if noise is None and self.top_k == 2:
                    # top2gating draws Gumbel(0, 1) noise for the second choice (gumbel_rsample), in eval mode too
                    noise = -torch.log(-torch.log(torch.rand((S, E), dtype=f32, device=dev).clamp_(1e-20, 1.0 - 1e-7)))
                elif noise is None and self.use_rts:
                    noise = torch.rand((S, E), dtype=f32, device=dev)
                route = ops.moe_route(n2, L.wg.detach(), self.top_k, C, noise)
                l_aux.append(route["l_aux"])
                gate_logits.append(route["logits"])
                rows = E * C
                xin = T.expert_buffer(rows, D, C, route["kept"], dev)  # (rows past kept[e]: zero; the rest: dispatch)
                ops.moe_dispatch(n2, route["slot"], rows, out=xin)
                sv.update(route=route, C=C)
            else:
                C, rows, xin, route = S, S, n2, None
            g = T.expert_buffer(rows, F_, C, route["kept"], dev) if route is not None else \
                torch.empty((rows, F_), dtype=bf16, device=dev)
            u = T.expert_buffer(rows, F_, C, route["kept"], dev) if route is not None else torch.empty_like(g)
            y = T.expert_buffer(rows, D, C, route["kept"], dev) if route is not None else None
            a_mlp = []
            # gate, up and SiLU(gate) * up in ONE dual GEMM per expert (g and u are kept for the backward) when no adapter of
            # the pair needs the epilogue / separate-pass route; else two GEMMs + the SiLU * up pass
            fuse_gu = LORA_FUSE and all(
                (not [lo for lo in (L.lo_mlp[e]["gate_proj"], L.lo_mlp[e]["up_proj"]) if lo is not None])
                or self._site_usable(L.ext_f[("gateup", e)], [lo for lo in (L.lo_mlp[e]["gate_proj"], L.lo_mlp[e]["up_proj"])
                                                              if lo is not None]) for e in range(E))
            h = None
            if fuse_gu:
                h = T.expert_buffer(rows, F_, C, route["kept"], dev) if route is not None else \
                    torch.empty((rows, F_), dtype=bf16, device=dev)
            for e in range(E):
                r0, r1 = e * C, (e + 1) * C
                md = route["kept"][e:e + 1] if route is not None else None
                force = "tc" if md is not None else None
                am = {}
                if fuse_gu:
                    site = L.ext_f[("gateup", e)]
                    ext = None
                    if site is not None:
                        buf = self._pad(r1 - r0, dev)
                        for nm in ("gate_proj", "up_proj"):
                            lo = L.lo_mlp[e][nm]
                            if lo is not None:
                                am[nm] = self._lora_a(lo, xin[r0:r1], pad=(buf, site.col[id(lo)]))
                        ext = (buf, site.b)
                    ops.linear(xin[r0:r1], L.w_gate[e], weight2=L.w_up[e], out=h[r0:r1], m_dev=md, force="tc", ext=ext,
                               dual_out=(g[r0:r1], u[r0:r1]))
                    a_mlp.append(am)
                    continue
                for nm, wt, buf in (("gate_proj", L.w_gate[e], g), ("up_proj", L.w_up[e], u)):
                    lo = L.lo_mlp[e][nm]
                    _, sav = self._linear_lora(xin[r0:r1], wt, [(0, lo)] if lo is not None else [], L.ext_f[(nm, e)],
                                               out=buf[r0:r1], m_dev=md, force=force)
                    if 0 in sav:
                        am[nm] = sav[0]
                a_mlp.append(am)
            if h is None:
                h = T.silu_mul(g, u)
            for e in range(E):
                r0, r1 = e * C, (e + 1) * C
                md = route["kept"][e:e + 1] if route is not None else None
                lo = L.lo_mlp[e]["down_proj"]
                one = [(0, lo)] if lo is not None else []
                if route is None:
                    x_next, sav = self._linear_lora(h, L.w_down[0], one, L.ext_f[("down_proj", 0)], residual=h1)
                else:
                    _, sav = self._linear_lora(h[r0:r1], L.w_down[e], one, L.ext_f[("down_proj", e)], out=y[r0:r1],
                                               m_dev=md, force="tc")
                if 0 in sav:
                    a_mlp[e]["down_proj"] = sav[0]
            if route is not None:
                x_next = ops.moe_combine(y, route["slot"], route["gate"], residual=h1)
            sv.update(xin=xin, g=g, u=u, h=h, y=y, a_mlp=a_mlp)
            saved.append(sv)
            if self.debug is not None:
                self.debug.update({f"f{li}.n1": n1, f"f{li}.o": o2, f"f{li}.h1": h1, f"f{li}.n2": n2,
                                   f"f{li}.out": x_next, f"f{li}.y": y, f"f{li}.slot": route["slot"] if route else None,
                                   f"f{li}.gate": route["gate"] if route else None})
            x = x_next
        hidden = ops.rmsnorm(x, self.norm_w, eps)
        la = torch.cat(l_aux) if l_aux else None
        return hidden, la, gate_logits, dict(layers=saved, x_last=x, B=B, T=Tn, kv_mask=km)

    # ------------------------------------------------------------------ backward
    def backward(self, saved, dhidden, aux_scale=0.0):
        """dhidden bf16 [S, D] (gradient of the loss w.r.t. the normed last hidden state). Accumulates every trainable
        parameter's gradient into the arena; returns dx0 bf16 [S, D] (gradient w.r.t. the input embeddings)."""
        D, H, hd, eps = self.D, self.H, self.hd, self.eps
        B, Tn, km = saved["B"], saved["T"], saved["kv_mask"]
        S = B * Tn
        dev = dhidden.device
        ar = self.arena
        cos, sin = self.cos, self.sin
        scale = 1.0 / math.sqrt(hd)
        dx = T.rmsnorm_bwd(saved["x_last"], self.norm_w, dhidden, eps, dweight=ar.of(self.norm_w))
        if self.reducer is not None and ar.of(self.norm_w) is not None:
            self.reducer.ready(ar.end_of(self.norm_w))
        for li in range(len(self.layers) - 1, -1, -1):
            L, sv = self.layers[li], saved["layers"][li]
            E, C, route = L.E, sv.get("C", S), sv.get("route")
            g, u, h, xin = sv["g"], sv["u"], sv["h"], sv["xin"]
            rows = g.shape[0]
            if route is not None:
                dy, dgate = T.moe_combine_bwd(dx, sv["y"], route["slot"], route["gate"], rows, C=C, kept=route["kept"])
            else:
                dy, dgate = dx, None
            # dh = dy W_down (+ adapter) is consumed by the SiLU(gate) * up backward in the SAME launch's epilogue (g <- dg,
            # u <- du in place) where every adapter of down_proj rides in the extension block; else dh goes through memory
            # (measured on B200: the fused epilogue costs the GEMMs 9 ms per step and saves a 3.8 ms pass -- two loads, an
            # exponential and two stores per element do not hide behind a K = 4096 main loop -- so it is off unless
            # MPL_SILU_BWD_FUSE=1)
            fuse_sb = SILU_BWD_FUSE and LORA_FUSE and all(L.lo_mlp[e]["down_proj"] is None or
                                        self._site_usable(L.ext_b[("down_proj", e)], [L.lo_mlp[e]["down_proj"]])
                                        for e in range(E))
            dh = None
            if not fuse_sb:
                dh = T.expert_buffer(rows, h.shape[1], C, route["kept"], dev) if route is not None else torch.empty_like(h)
            for e in range(E):
                r0, r1 = e * C, (e + 1) * C
                md = route["kept"][e:e + 1] if route is not None else None
                force = "tc" if md is not None else None
                lo = L.lo_mlp[e]["down_proj"]
                ad = [(lo, h[r0:r1], sv["a_mlp"][e]["down_proj"], dy[r0:r1])] if lo is not None else []
                if fuse_sb:
                    self._dgrad_lora(dy[r0:r1], L.w_downT[e], ad, L.ext_b[("down_proj", e)], m_dev=md, force="tc",
                                     silu_bwd=(g[r0:r1], u[r0:r1]))
                else:
                    self._dgrad_lora(dy[r0:r1], L.w_downT[e], ad, L.ext_b[("down_proj", e)], out=dh[r0:r1], m_dev=md,
                                     force=force)
            if not fuse_sb:
                T.silu_mul_bwd(g, u, dh)  # g <- dg, u <- du
            dxin = T.expert_buffer(rows, xin.shape[1], C, route["kept"], dev) if route is not None else torch.empty_like(xin)
            for e in range(E):
                r0, r1 = e * C, (e + 1) * C
                md = route["kept"][e:e + 1] if route is not None else None
                force = "tc" if md is not None else None
                for nm, buf, wT, res in (("gate_proj", g, L.w_gateT[e], None), ("up_proj", u, L.w_upT[e], dxin[r0:r1])):
                    lo = L.lo_mlp[e][nm]
                    ad = [(lo, xin[r0:r1], sv["a_mlp"][e][nm], buf[r0:r1])] if lo is not None else []
                    self._dgrad_lora(buf[r0:r1], wT, ad, L.ext_b[(nm, e)], out=dxin[r0:r1], residual=res, m_dev=md,
                                     force=force)
            if route is not None:
                ones = torch.ones_like(route["gate"])
                dn2 = ops.moe_combine(dxin, route["slot"], ones)
                dlogits = T.moe_router_bwd(route, dgate, L.wg.detach(), dn2, aux_scale=aux_scale)
                gwg = ar.of(L.wg)
                if gwg is not None:
                    T.rank_wgrad(sv["n2"], dlogits, gwg, transposed=True)
            else:
                dn2 = dxin
            dh1 = T.rmsnorm_bwd(sv["h1"], L.ln2, dn2, eps, add=dx, dweight=ar.of(L.ln2))
            # attention block
            o2 = sv["o"].view(S, D)
            ad = [(L.lo["o_proj"], o2, sv["a_o_proj"], dh1)] if L.lo["o_proj"] is not None else []
            do = self._dgrad_lora(dh1, L.woT, ad, L.ext_b["o"])
            qkv = sv["qkv"]
            q5 = qkv.view(B, Tn, 3, H, hd)
            dqkv = torch.empty_like(qkv)
            d5 = dqkv.view(B, Tn, 3, H, hd)
            dq32 = T.attention_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], sv["o"], do.view(B, Tn, H, hd), sv["lse"], scale,
                                   d5[:, :, 1], d5[:, :, 2], causal=True, kv_mask=km)
            T.rope_bwd(dq32, d5[:, :, 0], d5[:, :, 1], cos, sin, 0)
            ad = [(L.lo[nm], sv["n1"], sv["a_" + nm], dqkv[:, j * D:(j + 1) * D])
                  for j, nm in enumerate(("q_proj", "k_proj", "v_proj")) if L.lo[nm] is not None]
            dn1 = self._dgrad_lora(dqkv, L.wqkvT, ad, L.ext_b["qkv"])
            dx_in = dx
            dx = T.rmsnorm_bwd(sv["x"], L.ln1, dn1, eps, add=dh1, dweight=ar.of(L.ln1))
            if self.debug is not None:
                self.debug.update({f"b{li}.dout": dx_in, f"b{li}.dn2": dn2, f"b{li}.dh1": dh1, f"b{li}.do": do,
                                   f"b{li}.dqkv": dqkv, f"b{li}.dn1": dn1, f"b{li}.dx": dx, f"b{li}.dxin": dxin,
                                   f"b{li}.dy": dy, f"b{li}.dgate": dgate})
            saved["layers"][li] = None  # free this layer's activations
            if self.reducer is not None and L.arena_end is not None:
                self.reducer.ready(L.arena_end)
        return dx


# ------------------------------------------------------------------------------------------- autograd tape nodes
class _StackFn(torch.autograd.Function):
    """hidden = decoder_stack(embeds). `anchor` is a dummy leaf that makes the output require grad."""

    @staticmethod
    def forward(ctx, anchor, tr, embeds, B, Tn, kv_mask, moe_noise, splice_idx, region_ctx, proj_ctx, *params):
        # (params: every trainable parameter in arena order when tr.foreign_grads -- autograd then routes their
        # gradients, returned by backward(), through the ordinary AccumulateGrad nodes and their hooks)
        ctx.n_params = len(params)
        hidden, l_aux, gate_logits, saved = tr.stack.forward(embeds, B, Tn, kv_mask, moe_noise)
        ctx.tr, ctx.saved, ctx.splice_idx, ctx.region_ctx, ctx.proj_ctx = tr, saved, splice_idx, region_ctx, proj_ctx
        tr.last_gate_logits = gate_logits
        if l_aux is None:
            l_aux = torch.zeros(0, dtype=f32, device=embeds.device)
        return hidden.view(B, Tn, -1), l_aux

    @staticmethod
    def backward(ctx, dhidden, dl_aux):
        tr = ctx.tr
        D = ctx.saved["x_last"].shape[-1]
        aux = 0.0
        if dl_aux is not None and dl_aux.numel() > 0 and tr.stack.aux_coef != 0.0:
            aux = float(dl_aux[0])  # d loss / d l_aux (same for every layer); one host read per step, only with aux loss
        if dhidden is None:
            dh = torch.zeros((ctx.saved["B"] * ctx.saved["T"], D), dtype=bf16, device=tr.anchor.device)
        else:
            dh = dhidden.to(bf16).reshape(-1, D).contiguous()
        if tr.reducer is not None and tr.arena.of(tr.lm_head_weight) is not None:
            # every consumer of `hidden` (CE head, mask head) has finished its backward by the time this node runs: the
            # arena prefix up to and including lm_head is final whatever order autograd ran those branches in
            tr.reducer.ready(tr.arena.end_of(tr.lm_head_weight))
        dx0 = tr.stack.backward(ctx.saved, dh, aux_scale=aux)
        g_emb = tr.arena.of(tr.embed_weight)
        if g_emb is not None and ctx.splice_idx is not None:
            T.scatter_add_rows(dx0, ctx.splice_idx, dtable=g_emb)
            if tr.reducer is not None:
                tr.reducer.ready(tr.arena.end_of(tr.embed_weight))
        if ctx.region_ctx is not None:
            # region_fea_adapter (medplib_arch.py:131,208,580-614): feature = sampled_raw_clip W^T + b, so
            # dW += dfeat^T sampled_raw, db += sum dfeat, with dfeat = the input-embedding gradient at the region slots
            ad = tr.model.model.region_fea_adapter
            gW, gb = tr.arena.of(ad.weight), tr.arena.of(ad.bias)
            dfeat = ops.gather_rows(ctx.region_ctx["pos"], table=dx0)
            if gW is not None:
                T.gemm_small(dfeat, ctx.region_ctx["sampled"], out=gW, trans_a=True, accumulate=True)
            if gb is not None:
                T.col_sum(dfeat, gb)
        if ctx.proj_ctx is not None:
            _vision_backward(tr, ctx.proj_ctx, dx0)
        ctx.saved = None
        tr.micro_steps += 1  # the stack's backward is the last node of a micro-step
        grads = ()
        if ctx.n_params:
            # foreign-optimizer mode: this node runs last (every other tape node feeds from its output), so the arena is
            # complete -- hand each parameter its gradient in the parameter's dtype and start the next micro-step from
            # zero (accumulation is the foreign engine's business, as are the all-reduce and the optimizer step)
            grads = tuple(tr.arena.of(p).to(p.dtype, copy=True) for p in tr.arena.params)  # (fp32 params: a COPY, not a view)
            tr.arena.zero_()
            tr.micro_steps = 0
        return (torch.zeros(1, dtype=f32, device=dx0.device),) + (None,) * 9 + grads


def _wgrad_tc(dy, x, out, accumulate):
    """out (f32 [N, K]) (+)= dy^T x on the tensor cores: both operands transposed to K-major over the rows."""
    M = dy.shape[0]
    Mp = (M + 63) // 64 * 64  # whole K blocks of the tcgen05 tile; pad columns are zero
    dyT = T.transpose(dy, ld_out=Mp)[:dy.shape[1]]
    xT = T.transpose(x, ld_out=Mp)[:x.shape[1]]
    if not accumulate:
        ops.linear(dyT, xT, out_dtype=f32, out=out, force="tc")
    else:
        tmp = ops.linear(dyT, xT, out_dtype=f32, force="tc")
        T.col_sum(tmp.view(1, -1), out.view(-1))


def _projector_backward(tr, x, dF):
    """mm_projector = Linear(1024, D) + GELU + Linear(D, D) (multimodal_projector/builder.py:39-46) on the frozen CLIP
    features x: weight / bias gradients from dF = the gradient at the projector's output rows (no gradient into CLIP)."""
    proj = tr.model.model.mm_projector
    ar = tr.arena
    acc = tr.micro_steps > 0
    z0 = ops.linear(x, proj[0].weight.detach(), bias=proj[0].bias.detach())
    h1 = T.act_fwd(z0, "gelu")
    if ar.of(proj[2].weight) is not None:
        _wgrad_tc(dF, h1, ar.of(proj[2].weight), acc)
    if ar.of(proj[2].bias) is not None:
        T.col_sum(dF, ar.of(proj[2].bias))
    dh1 = ops.linear(dF, T.transpose(proj[2].weight.detach()))
    dz0 = T.act_bwd(z0, dh1, "gelu")
    if ar.of(proj[0].weight) is not None:
        _wgrad_tc(dz0, x, ar.of(proj[0].weight), acc)
    if ar.of(proj[0].bias) is not None:
        T.col_sum(dz0, ar.of(proj[0].bias))


def _vision_backward(tr, pc, dx0):
    """Backward of the vision-side adapters from dx0, the gradient w.r.t. the spliced input embeddings:
    TokenCompressor (AdaptiveAvgPool1d 576 -> 256 + LayerNorm + Linear, medplib_arch.py:67-77) and, below it, the
    mm_projector; MaskTokenEncoder (4 x conv3x3 s2 + GELU, pool, Linear, LayerNorm, medplib_arch.py:80-108). The forward
    intermediates are recomputed here (the adapters are < 0.1 % of a step); the CLIP tower is frozen."""
    mdl = tr.model.model
    ar = tr.arena
    acc = tr.micro_steps > 0
    x = pc["feats"].contiguous()
    dF = ops.gather_rows(pc["pos"], table=dx0)  # [image feature rows, D]; rows the prompt never used are zero
    if pc.get("compress"):
        c = mdl.mm_token_compressor
        proj = mdl.mm_projector
        if pc.get("comp_train") or pc.get("proj_train"):
            n_img = pc["n_img"]
            if isinstance(proj, nn.Linear):
                x1 = ops.linear(x, proj.weight.detach(), bias=proj.bias.detach())
            else:
                x1 = ops.linear(x, proj[0].weight.detach(), bias=proj[0].bias.detach(), act="gelu")
                x1 = ops.linear(x1, proj[2].weight.detach(), bias=proj[2].bias.detach())
            D = x1.shape[-1]
            t_in = x1.shape[0] // n_img
            pooled = T.token_pool(x1.view(n_img, t_in, D), c.num_tokens).view(-1, D)
            ln_out = ops.layernorm(pooled, c.norm.weight.detach(), c.norm.bias.detach(), c.norm.eps)
            if ar.of(c.proj.weight) is not None:
                _wgrad_tc(dF, ln_out, ar.of(c.proj.weight), acc)
            if ar.of(c.proj.bias) is not None:
                T.col_sum(dF, ar.of(c.proj.bias))
            d_ln = ops.linear(dF, T.transpose(c.proj.weight.detach()))
            d_pooled = T.layernorm_bwd(pooled, c.norm.weight.detach(), d_ln, c.norm.eps, dweight=ar.of(c.norm.weight),
                                       dbias=ar.of(c.norm.bias))
            dF = T.token_pool_bwd(d_pooled.view(n_img, c.num_tokens, D), t_in).view(-1, D) if pc.get("proj_train") else None
    if pc.get("proj_train") and dF is not None:
        _projector_backward(tr, x, dF)
    if pc.get("mask") is not None:
        _mask_encoder_backward(tr, pc["mask"], dx0)


def _mask_encoder_backward(tr, mc, dx0):
    """MaskTokenEncoder backward (medplib_arch.py:80-108): the convs are im2col GEMMs on NHWC rows, so their dgrad is a
    GEMM + col2im (mpl_col2im_nhwc) and their wgrad dz^T cols."""
    me = tr.model.model.mask_encoder
    ar = tr.arena
    x = mc["images"]
    if x.dim() == 3:
        x = x.unsqueeze(1)
    x = x[:, :1].to(bf16).permute(0, 2, 3, 1).contiguous()
    n = x.shape[0]
    x = torch.cat([x, x.new_zeros(*x.shape[:3], 7)], dim=-1)  # C = 1 -> 8 (16-byte rows), as the forward
    saved = []
    for i in (0, 2, 4, 6):
        conv = me.encoder[i]
        H, Cin = x.shape[1], x.shape[-1]
        w = conv.weight.detach()
        if w.shape[1] != Cin:
            w = torch.cat([w, w.new_zeros(w.shape[0], Cin - w.shape[1], 3, 3)], dim=1)
        w2 = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()
        cols = ops.im2col_nhwc(x, 3, 3, 2, 1)
        Ho = (H + 2 - 3) // 2 + 1
        z = ops.linear(cols, w2, bias=conv.bias.detach())
        saved.append((conv, H, Cin, w2, cols, z))
        x = T.act_fwd(z, "gelu").view(n, Ho, Ho, -1)
    C = x.shape[-1]
    feats = x.reshape(n, -1, C)
    pooled = T.token_pool(feats, me.num_tokens).view(-1, C)
    y = ops.linear(pooled, me.proj.weight.detach(), bias=me.proj.bias.detach())
    d_out = ops.gather_rows(mc["pos"], table=dx0)  # [n * num_tokens, D]
    dy = T.layernorm_bwd(y, me.norm.weight.detach(), d_out, me.norm.eps, dweight=ar.of(me.norm.weight),
                         dbias=ar.of(me.norm.bias))
    if ar.of(me.proj.weight) is not None:
        T.gemm_small(dy, pooled, out=ar.of(me.proj.weight), trans_a=True, accumulate=True)
    if ar.of(me.proj.bias) is not None:
        T.col_sum(dy, ar.of(me.proj.bias))
    d_pooled = ops.linear(dy, T.transpose(me.proj.weight.detach()))
    d = T.token_pool_bwd(d_pooled.view(n, me.num_tokens, C), feats.shape[1]).view(-1, C)
    for li in range(len(saved) - 1, -1, -1):
        conv, H, Cin, w2, cols, z = saved[li]
        dz = T.act_bwd(z, d, "gelu")
        gW = ar.of(conv.weight)
        if gW is not None:
            # dW2 [Cout, (ky, kx, ci)] = dz^T cols; the parameter is [Cout, Cin, ky, kx] (first conv: channel 0 of the 8)
            dW2 = torch.zeros((dz.shape[1], cols.shape[1]), dtype=f32, device=dz.device)
            T.gemm_small(dz, cols, out=dW2, trans_a=True, accumulate=True)  # (accumulate mode: split-K over the rows)
            cin_p = conv.weight.shape[1]
            gW.add_(dW2.view(dW2.shape[0], 3, 3, Cin)[..., :cin_p].permute(0, 3, 1, 2))
        if ar.of(conv.bias) is not None:
            T.col_sum(dz, ar.of(conv.bias))
        if li > 0:
            dcols = ops.linear(dz, T.transpose(w2))
            d = T.col2im_nhwc(dcols, n, H, H, Cin, 3, 3, 2, 1).view(-1, Cin)


class _HeadCEFn(torch.autograd.Function):
    """medplib_moe_llama.py:381-421: fp32 logits = lm_head(hidden); shifted cross-entropy over the valid labels."""

    @staticmethod
    def forward(ctx, hidden, tr, labels):
        Bn, Tn, D = hidden.shape
        h2 = hidden.reshape(-1, D)
        W = tr.lm_head_weight
        logits = ops.linear(h2, W.detach(), out_dtype=f32)
        shift = torch.full_like(labels, -100)
        shift[:, :-1] = labels[:, 1:]
        lab = shift.reshape(-1).contiguous()
        lse, acc = T.ce_fwd(logits, lab)
        ctx.tr, ctx.h2, ctx.logits, ctx.lab, ctx.lse, ctx.acc = tr, h2, logits, lab, lse, acc
        ctx.shape = hidden.shape
        loss = acc[0] / acc[1]
        logits3 = logits.view(Bn, Tn, -1)
        ctx.mark_non_differentiable(logits3)
        return loss, logits3

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        tr, h2, logits = ctx.tr, ctx.h2, ctx.logits
        W = tr.lm_head_weight
        V, D = W.shape
        S = h2.shape[0]
        Vp = (V + 7) // 8 * 8
        dl = T.ce_bwd(logits, ctx.lab, ctx.lse, ctx.acc, grad_out=dloss.reshape(1).float().contiguous(), ldd=Vp)
        WT = T.transpose(W.detach(), ld_out=Vp)  # [D, Vp], pad columns zero
        dh = ops.linear(dl, WT)
        gW = tr.arena.of(W)
        if gW is not None:
            Sp = (S + 7) // 8 * 8
            dlT = T.transpose(dl, ld_out=Sp)[:V]  # [V, Sp]
            hT = T.transpose(h2, ld_out=Sp)  # [D, Sp]
            if tr.micro_steps == 0:
                ops.linear(dlT, hT, out_dtype=f32, out=gW, force="tc")  # first micro-step: write straight into the arena
            else:  # gradient accumulation: the GEMM overwrites, so go through a temporary and add
                tmp = ops.linear(dlT, hT, out_dtype=f32, force="tc")
                T.col_sum(tmp.view(1, -1), gW.view(-1))
            # (no reducer.ready() here: the regions in front of lm_head in the arena -- mask decoder, text_hidden_fcs --
            # are written by other tape nodes whose order relative to this one autograd does not promise; the stack's
            # backward, which depends on every branch, releases them)
        ctx.logits = ctx.h2 = None
        return dh.view(ctx.shape), None, None


class Trainer:
    """Owns the arena, the reducer, the optimizer and the tape nodes for one MedPLIBForCausalLM. Created lazily by
    the model on its first training forward (model.trainer()), or explicitly to pick hyper-parameters:

        tr = model.trainer(lr=3e-4)          # after attach_lora / set_trainable, model in bf16 on the GPU
        out = model(**batch); out["loss"].backward(); tr.step()
    """

    def __init__(self, model, lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0,
                 bucket_elems=64 << 20, group=None, foreign_grads=False):
        """foreign_grads=True: for a foreign engine / optimizer (real DeepSpeed ZeRO, torch.optim, DDP): after
        ``loss.backward()`` every trainable parameter has an ordinary ``.grad`` (parameter dtype) delivered through
        autograd -- gradient hooks fire, ``.grad`` accumulates across backward calls -- and this Trainer neither
        all-reduces nor steps. Costs one cast pass over the arena per backward (318 M parameters: ~0.4 ms)."""
        model._check_ready()
        self.model = model
        self.foreign_grads = bool(foreign_grads)
        dev = model.lm_head.weight.device
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise _lib.MplError("no trainable parameters (call attach_lora / set_trainable first)")
        self.arena = GradArena(self._order(model, named), dev)
        self.reducer = BucketReducer(self.arena, bucket_elems, group)
        self.reducer.owner = self
        self.stack = LlamaTrainStack(model, self.arena, self.reducer)
        self.opt = FusedAdamW(self.arena, lr, betas, eps, weight_decay, max_grad_norm)
        self.lm_head_weight = model.lm_head.weight
        self.embed_weight = model.model.embed_tokens.weight
        self.anchor = torch.zeros(1, dtype=f32, device=dev, requires_grad=True)
        self.loss_scale = 1.0
        self.last_gate_logits = None
        self.micro_steps = 0  # backward passes accumulated in the arena since the last step()
        if self.foreign_grads:
            self.reducer.on = False
        self._sig = self.signature(model)

    @staticmethod
    def signature(model):
        return tuple((n, p.data_ptr()) for n, p in model.named_parameters() if p.requires_grad)

    @staticmethod
    def _order(model, named):
        """Backward-completion order: mask head + text_hidden_fcs + lm_head + final norm, decoder layers last-to-first,
        then embed_tokens and everything else (region adapter, projector-side modules)."""
        def key(item):
            n = item[0]
            if "visual_model" in n or "text_hidden_fcs" in n:
                return (0, 0)
            if n.startswith("lm_head"):
                return (1, 0)
            if n == "model.norm.weight":
                return (2, 0)
            if n.startswith("model.layers."):
                return (3, -int(n.split(".")[2]))
            if "embed_tokens" in n:
                return (4, 0)
            return (5, 0)
        return sorted(named, key=key)

    # tape
    def stack_hidden(self, embeds, kv_mask=None, moe_noise=None, splice_idx=None, region_ctx=None, proj_ctx=None):
        self.stack.training = bool(self.model.training)
        B, Tn, D = embeds.shape
        x = embeds.to(bf16).reshape(B * Tn, D).contiguous()
        extra = ()
        if self.foreign_grads:
            self.model.refresh_trained()  # the foreign optimizer may have moved the weights since the last forward
            extra = tuple(self.arena.params)
        return _StackFn.apply(self.anchor, self, x, B, Tn, kv_mask, moe_noise, splice_idx, region_ctx, proj_ctx, *extra)

    def head_ce(self, hidden, labels):
        return _HeadCEFn.apply(hidden, self, labels)

    def _accum_ok(self):
        # buckets may only have been launched during the LAST backward (reducer.reset() happens in zero_grad)
        return getattr(self.reducer, "launch_micro_step", self.micro_steps - 1) == self.micro_steps - 1

    def zero_grad(self):
        self.arena.zero_()
        self.reducer.reset()
        self.micro_steps = 0

    def no_sync(self):
        """Gradient accumulation (DeepSpeed gradient_accumulation_steps / DDP.no_sync): backward passes inside this
        context only add into the arena; the buckets are all-reduced during the first backward OUTSIDE it (or by
        step()), so every bucket is exchanged exactly once per optimizer step. Scale the loss by 1/k yourself."""
        tr = self

        class _NoSync:
            def __enter__(self_):
                self_.prev = tr.reducer.on
                tr.reducer.on = False

            def __exit__(self_, *a):
                tr.reducer.on = self_.prev

        return _NoSync()

    def step(self, lr=None):
        """all-reduce what is left (mean over the data-parallel group), clip, AdamW, zero the arena."""
        if self.foreign_grads:
            raise _lib.MplError("this Trainer was built with foreign_grads=True: the gradients are in p.grad and the "
                                "optimizer step belongs to the foreign engine")
        if self.micro_steps > 1 and self.reducer.on and self.reducer.next > 0 and not self._accum_ok():
            raise _lib.MplError("backward() ran more than once since the last step() with the bucket all-reduce enabled: "
                                "wrap all but the last micro-step in `with trainer.no_sync():`")
        scale = self.reducer.finish()
        self.opt.step(grad_scale=scale, lr=lr)
        self.zero_grad()
        self.model.refresh_trained()


def merge_lora(model):
    """Offline tool (what peft's merge_and_unload does before the reference evaluates a trained model): fold every
    adapter into its base weight, W += scaling * B A, and remove the adapter modules. Not on the hot path."""
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, nn.Linear) and hasattr(mod, "lora_A"):
                A, Bm = mod.lora_A["default"].weight, mod.lora_B["default"].weight
                mod.weight.add_((Bm.float() @ A.float() * mod.scaling["default"]).to(mod.weight.dtype))
                del mod.lora_A, mod.lora_B, mod.scaling, mod.r
    if hasattr(model, "refresh_engines"):
        model.refresh_engines()
    model._trainer = None
    return model
