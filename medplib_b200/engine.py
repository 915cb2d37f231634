"""Host side of the native stack runners (mpl_llama_forward / mpl_clip_forward / mpl_sam_*): builds the C weight
tables from tensors named like the reference's parameters and owns the device workspaces.

An engine never copies a trainable weight: the tables hold ``data_ptr()`` of the caller's tensors (nn.Parameters read
in place). Only frozen convolution weights are repacked once into GEMM layout (SAM adapters / neck / upscaler, CLIP
patch embedding); ``refresh()`` rebuilds tables and repacks after weights were replaced (load_state_dict, .to()).
"""
import ctypes
import math

import torch

from . import _lib, ops

bf16 = torch.bfloat16


_SINK = None  # list collecting every tensor whose address goes into a weight table (keeps the storage alive)


def _p(t):
    if t is None:
        return None
    if _SINK is not None:
        _SINK.append(t)
    return t.data_ptr()


class _collect:
    """with _collect(engine): every _p() inside records its tensor in engine._refs."""

    def __init__(self, owner):
        self.owner = owner

    def __enter__(self):
        global _SINK
        self.owner._refs = []
        self.prev, _SINK = _SINK, self.owner._refs

    def __exit__(self, *a):
        global _SINK
        _SINK = self.prev


def _vp(t):
    """c_void_p argument for a direct call (a bare Python int would be truncated to a C int)."""
    return ctypes.c_void_p(t.data_ptr() if t is not None else None)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Workspace:
    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


def rope_tables(head_dim, max_pos, theta, device):
    """HF-4.31 LlamaRotaryEmbedding cache: fp32 angles, tables cast to bf16."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    t = torch.arange(max_pos, dtype=torch.float32)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(bf16).to(device).contiguous(), emb.sin().to(bf16).to(device).contiguous()


class KVCache:
    """bf16 [L, B, H, Tmax, d] keys and values (HF tuple layout per layer is a view of this)."""

    def __init__(self, n_layers, B, H, Tmax, d, device):
        self.k = torch.empty((n_layers, B, H, Tmax, d), dtype=bf16, device=device)
        self.v = torch.empty((n_layers, B, H, Tmax, d), dtype=bf16, device=device)
        self.B, self.Tmax, self.len = B, Tmax, 0


class LlamaEngine:
    """Runs the decoder stack of a state dict with the reference's names (prefix e.g. 'model.')."""

    def __init__(self, sd, cfg, prefix="model."):
        self.cfg = dict(cfg)
        self.prefix = prefix
        self.ws = _Workspace()
        self.refresh(sd)

    def refresh(self, sd):
        with _collect(self):
            self._refresh(sd)

    def _refresh(self, sd):
        c, p = self.cfg, self.prefix
        L = c["num_layers"]
        self._keep = []
        layers = (_lib.LlamaLayer * L)()
        dev = None
        for i in range(L):
            lp = f"{p}layers.{i}."
            lay = layers[i]
            lay.input_ln = _p(sd[lp + "input_layernorm.weight"])
            lay.wq, lay.wk = _p(sd[lp + "self_attn.q_proj.weight"]), _p(sd[lp + "self_attn.k_proj.weight"])
            lay.wv, lay.wo = _p(sd[lp + "self_attn.v_proj.weight"]), _p(sd[lp + "self_attn.o_proj.weight"])
            lay.post_ln = _p(sd[lp + "post_attention_layernorm.weight"])
            dev = sd[lp + "input_layernorm.weight"].device
            wgk = lp + "mlp.deepspeed_moe.gate.wg.weight"
            if wgk in sd:
                wg = sd[wgk]
                if wg.dtype != torch.float32:
                    raise _lib.MplError("the MoE gate (wg) must stay fp32 like DeepSpeed's TopKGate")
                E = wg.shape[0]
                lay.wg, lay.n_experts = _p(wg), E
                for e in range(E):
                    ep = f"{lp}mlp.deepspeed_moe.experts.deepspeed_experts.{e}."
                    lay.w_gate[e], lay.w_up[e] = _p(sd[ep + "gate_proj.weight"]), _p(sd[ep + "up_proj.weight"])
                    lay.w_down[e] = _p(sd[ep + "down_proj.weight"])
            else:
                lay.wg, lay.n_experts = None, 1
                lay.w_gate[0], lay.w_up[0] = _p(sd[lp + "mlp.gate_proj.weight"]), _p(sd[lp + "mlp.up_proj.weight"])
                lay.w_down[0] = _p(sd[lp + "mlp.down_proj.weight"])
        self.layers = layers
        self.device = dev
        hd = c["hidden_size"] // c["num_heads"]
        self.rope_len = int(c.get("max_position_embeddings", 4096))
        self.cos, self.sin = rope_tables(hd, self.rope_len, c.get("rope_theta", 1e4), dev)
        m = _lib.LlamaModel()
        m.n_layers, m.hidden, m.n_heads, m.ffn = L, c["hidden_size"], c["num_heads"], c["intermediate_size"]
        m.rms_eps = c["rms_norm_eps"]
        moe = c.get("moe") or {}
        m.top_k = int(moe.get("top_k_experts", 1) or 1)
        m.min_capacity = int(moe.get("min_capacity", 0) or 0)
        m.layers = layers
        m.final_norm = _p(sd[p + "norm.weight"])
        m.rope_cos, m.rope_sin, m.rope_len = _p(self.cos), _p(self.sin), self.rope_len
        self.model = m
        self.head_dim = hd
        self.E = max([layers[i].n_experts for i in range(L)])
        self.attn_scratch = torch.zeros(5 << 20, dtype=torch.uint8, device=dev)  # split-K decode attention
        # decode plan: TMA descriptors of every weight matrix for the one-kernel decode step (llama_decode.cu)
        self.decode_plan = None
        self.use_decode_kernel = True
        if dev is not None and dev.type == "cuda":
            lib = _lib.load()
            nbytes = lib.mpl_llama_decode_plan_bytes(ctypes.byref(m))
            plan = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
            if lib.mpl_llama_decode_plan_build(ctypes.byref(m), _vp(plan), _stream()) == 0:
                self.decode_plan = plan

    def new_cache(self, B, Tmax):
        c = self.cfg
        return KVCache(c["num_layers"], B, c["num_heads"], Tmax, self.head_dim, self.device)

    def forward(self, x, cache, training=False, kv_mask=None, want_hidden_states=False, moe_noise=None,
                want_router=False, pos_dev=None, tk_dev=None, rope_pos=None):
        """x bf16 [B,T,D] inputs_embeds (clobbered: holds the last layer's output afterwards). Appends to cache.

        Returns dict(last_hidden_state [B,T,D], hidden_states (tuple or None), gate_logits [L,S,E] f32, l_aux [L],
        exp_counts [L,E])."""
        lib = _lib.load()
        c = self.cfg
        assert x.dtype == bf16 and x.is_cuda and x.is_contiguous() and x.dim() == 3
        B, T, D = x.shape
        L = c["num_layers"]
        moe = c.get("moe") or {}
        cf = moe.get("capacity_factor", 1.0) if training else moe.get("eval_capacity_factor", 1.0)
        self.model.capacity_factor = float(cf if cf is not None else 1.0)
        if cache.len + T > self.rope_len:
            self.rope_len = max(2 * self.rope_len, cache.len + T)
            self.cos, self.sin = rope_tables(self.head_dim, self.rope_len, c.get("rope_theta", 1e4), self.device)
            self.model.rope_cos, self.model.rope_sin, self.model.rope_len = _p(self.cos), _p(self.sin), self.rope_len
        nbytes = lib.mpl_llama_workspace_bytes(ctypes.byref(self.model), B, T)
        ws = self.ws.get(nbytes, x.device)
        out_norm = torch.empty_like(x)
        io = _lib.LlamaIO()
        io.x, io.out_norm = x.data_ptr(), out_norm.data_ptr()
        hs = None
        if want_hidden_states:
            hs = [torch.empty_like(x) for _ in range(L)]
            arr = (ctypes.c_void_p * L)(*[h.data_ptr() for h in hs])
            io.hidden_states = arr
        io.B, io.T, io.past_len = B, T, cache.len
        io.k_cache, io.v_cache, io.Tmax = cache.k.data_ptr(), cache.v.data_ptr(), cache.Tmax
        assert cache.B == B
        if kv_mask is not None:
            kv_mask = kv_mask.to(torch.uint8).contiguous()
            assert kv_mask.shape[0] == B and kv_mask.shape[1] >= cache.len + T
            io.kv_mask, io.kv_mask_stride = kv_mask.data_ptr(), kv_mask.stride(0)
        if pos_dev is not None:
            io.pos_dev, io.tk_dev = pos_dev.data_ptr(), tk_dev.data_ptr()
        if rope_pos is not None:  # continuous batching: per-sequence RoPE positions, shared cache column (decode step)
            assert T == 1 and rope_pos.dtype == torch.int32 and rope_pos.numel() == B and rope_pos.is_cuda
            io.rope_pos = rope_pos.data_ptr()
        if moe_noise is not None:
            noise = [n.contiguous() if n is not None else None for n in moe_noise]
            io.moe_noise = (ctypes.c_void_p * L)(*[_p(n) for n in noise])
        S = B * T
        gate_logits = l_aux = exp_counts = None
        if want_router and self.E > 1:
            gate_logits = torch.zeros((L, S, self.E), dtype=torch.float32, device=x.device)
            l_aux = torch.zeros((L,), dtype=torch.float32, device=x.device)
            exp_counts = torch.zeros((L, self.E), dtype=torch.int32, device=x.device)
            io.gate_logits, io.l_aux, io.exp_counts = gate_logits.data_ptr(), l_aux.data_ptr(), exp_counts.data_ptr()
        io.attn_scratch, io.attn_scratch_bytes = self.attn_scratch.data_ptr(), self.attn_scratch.numel()
        if self.decode_plan is not None and self.use_decode_kernel:
            io.decode_plan = self.decode_plan.data_ptr()
        io.workspace, io.workspace_bytes = ws.data_ptr(), ws.numel()
        _lib.check(lib.mpl_llama_forward(ctypes.byref(self.model), ctypes.byref(io), _stream()), "mpl_llama_forward")
        if pos_dev is None:
            cache.len += T
        hidden = tuple(hs) + (out_norm,) if hs is not None else None
        return dict(last_hidden_state=out_norm, hidden_states=hidden, gate_logits=gate_logits, l_aux=l_aux,
                    exp_counts=exp_counts)


class ClipEngine:
    """CLIP vision tower -> hidden_states[select_layer][:, 1:] (prefix e.g. 'model.vision_tower.vision_tower.')."""

    def __init__(self, sd, cfg, prefix, select_layer=-2):
        self.cfg, self.prefix, self.select_layer = dict(cfg), prefix, select_layer
        self.ws = _Workspace()
        self.refresh(sd)

    def refresh(self, sd):
        with _collect(self):
            self._refresh(sd)

    def _refresh(self, sd):
        c, p = self.cfg, self.prefix + "vision_model."
        L = c["num_layers"]
        n_run = self.select_layer if self.select_layer >= 0 else L + 1 + self.select_layer
        D, P = c["hidden_size"], c["patch_size"]
        k = 3 * P * P
        k_pad = (k + 7) // 8 * 8
        pw = sd[p + "embeddings.patch_embedding.weight"]
        self.patch_w = torch.zeros((D, k_pad), dtype=bf16, device=pw.device)
        self.patch_w[:, :k] = pw.reshape(D, k).to(bf16)
        layers = (_lib.ClipLayer * max(n_run, 1))()
        for i in range(n_run):
            lp = f"{p}encoder.layers.{i}."
            lay = layers[i]
            lay.ln1_w, lay.ln1_b = _p(sd[lp + "layer_norm1.weight"]), _p(sd[lp + "layer_norm1.bias"])
            lay.ln2_w, lay.ln2_b = _p(sd[lp + "layer_norm2.weight"]), _p(sd[lp + "layer_norm2.bias"])
            for nm, w, b in (("q_proj", "wq", "bq"), ("k_proj", "wk", "bk"), ("v_proj", "wv", "bv"),
                             ("out_proj", "wo", "bo")):
                setattr(lay, w, _p(sd[f"{lp}self_attn.{nm}.weight"]))
                setattr(lay, b, _p(sd[f"{lp}self_attn.{nm}.bias"]))
            lay.fc1_w, lay.fc1_b = _p(sd[lp + "mlp.fc1.weight"]), _p(sd[lp + "mlp.fc1.bias"])
            lay.fc2_w, lay.fc2_b = _p(sd[lp + "mlp.fc2.weight"]), _p(sd[lp + "mlp.fc2.bias"])
        self.layers = layers
        m = _lib.ClipModel()
        m.n_layers, m.hidden, m.n_heads, m.mlp = n_run, D, c["num_heads"], c["intermediate_size"]
        m.image_size, m.patch, m.k_pad, m.ln_eps = c["image_size"], P, k_pad, c.get("layer_norm_eps", 1e-5)
        m.patch_w = _p(self.patch_w)
        m.cls = _p(sd[p + "embeddings.class_embedding"])
        m.pos = _p(sd[p + "embeddings.position_embedding.weight"])
        m.pre_ln_w, m.pre_ln_b = _p(sd[p + "pre_layrnorm.weight"]), _p(sd[p + "pre_layrnorm.bias"])
        m.layers = layers
        self.model = m
        self.n_patches = (c["image_size"] // P) ** 2

    def forward(self, images):
        """images [B,3,S,S] (any float dtype) -> bf16 [B, n_patches, hidden]."""
        lib = _lib.load()
        images = images.to(bf16).contiguous()
        B = images.shape[0]
        ws = self.ws.get(lib.mpl_clip_workspace_bytes(ctypes.byref(self.model), B), images.device)
        out = torch.empty((B, self.n_patches, self.cfg["hidden_size"]), dtype=bf16, device=images.device)
        _lib.check(lib.mpl_clip_forward(ctypes.byref(self.model), _vp(images), B, _vp(out), _vp(ws),
                                        ctypes.c_longlong(ws.numel()), _stream()), "mpl_clip_forward")
        return out


def _window_maps(B, g, ws, device):
    """Index maps of window_partition / window_unpartition (image_encoder.py:299-345) for mpl_gather_rows."""
    gp = (g + ws - 1) // ws * ws
    nw = gp // ws
    tok = torch.full((gp, gp), -1, dtype=torch.int64)
    tok[:g, :g] = torch.arange(g * g).view(g, g)
    part = tok.view(nw, ws, nw, ws).permute(0, 2, 1, 3).reshape(nw * nw, ws * ws)  # [window, pos] -> token or -1
    parts, unparts = [], []
    for b in range(B):
        pb = part.clone()
        pb[pb >= 0] += b * g * g
        parts.append(pb.reshape(-1))
        un = torch.empty(g * g, dtype=torch.int64)
        flat = part.reshape(-1)
        rows = torch.arange(flat.numel())
        un[flat[flat >= 0]] = rows[flat >= 0] + b * nw * nw * ws * ws
        unparts.append(un)
    return (torch.cat(parts).to(torch.int32).to(device), torch.cat(unparts).to(torch.int32).to(device), nw * nw)


class SamEncoderEngine:
    """SAM-Med2D image encoder (prefix e.g. 'model.visual_model.image_encoder.')."""
    GLOBAL = (2, 5, 8, 11)
    WINDOW = 14

    def __init__(self, sd, cfg, prefix):
        self.cfg, self.prefix = dict(cfg), prefix
        self.ws = _Workspace()
        self._maps = {}
        self.refresh(sd)

    def refresh(self, sd):
        with _collect(self):
            self._refresh(sd)

    def _refresh(self, sd):
        c, p = self.cfg, self.prefix
        depth, D = c["depth"], c["embed_dim"]
        self._keep = []

        def keep(t):
            t = t.to(bf16).contiguous()
            self._keep.append(t)
            return t.data_ptr()

        blocks = (_lib.SamBlock * depth)()
        for i in range(depth):
            bp = f"{p}blocks.{i}."
            b = blocks[i]
            b.ln1_w, b.ln1_b = _p(sd[bp + "norm1.weight"]), _p(sd[bp + "norm1.bias"])
            b.ln2_w, b.ln2_b = _p(sd[bp + "norm2.weight"]), _p(sd[bp + "norm2.bias"])
            b.qkv_w, b.qkv_b = _p(sd[bp + "attn.qkv.weight"]), _p(sd[bp + "attn.qkv.bias"])
            b.proj_w, b.proj_b = _p(sd[bp + "attn.proj.weight"]), _p(sd[bp + "attn.proj.bias"])
            b.rel_pos_h, b.rel_pos_w = _p(sd[bp + "attn.rel_pos_h"]), _p(sd[bp + "attn.rel_pos_w"])
            b.window = 0 if i in self.GLOBAL else self.WINDOW
            b.lin1_w, b.lin1_b = _p(sd[bp + "mlp.lin1.weight"]), _p(sd[bp + "mlp.lin1.bias"])
            b.lin2_w, b.lin2_b = _p(sd[bp + "mlp.lin2.weight"]), _p(sd[bp + "mlp.lin2.bias"])
            if bp + "Adapter.norm.weight" in sd:
                b.ad_ch0, b.ad_ch2 = _p(sd[bp + "Adapter.channel.0.weight"]), _p(sd[bp + "Adapter.channel.2.weight"])
                b.ad_conv = keep(sd[bp + "Adapter.spatial.0.weight"].permute(0, 2, 3, 1).reshape(D, 9 * D))
                b.ad_convt = keep(sd[bp + "Adapter.spatial.2.weight"].permute(2, 3, 1, 0).reshape(16 * D, D))
                b.ad_norm_w, b.ad_norm_b = _p(sd[bp + "Adapter.norm.weight"]), _p(sd[bp + "Adapter.norm.bias"])
        self.blocks = blocks
        O = c["out_chans"]
        m = _lib.SamEncoder()
        m.depth, m.hidden, m.n_heads, m.mlp = depth, D, c["num_heads"], int(D * c.get("mlp_ratio", 4))
        m.image_size, m.patch, m.out_chans = c["image_size"], c["patch_size"], O
        m.patch_w = keep(sd[p + "patch_embed.proj.weight"].reshape(D, -1))
        m.patch_b = _p(sd[p + "patch_embed.proj.bias"])
        m.pos_embed = keep(sd[p + "pos_embed"].reshape(-1, D))
        m.blocks = blocks
        m.neck0_w = keep(sd[p + "neck.0.weight"].reshape(O, D))
        m.neck1_w, m.neck1_b = _p(sd[p + "neck.1.weight"]), _p(sd[p + "neck.1.bias"])
        m.neck2_w = keep(sd[p + "neck.2.weight"].permute(0, 2, 3, 1).reshape(O, 9 * O))
        m.neck3_w, m.neck3_b = _p(sd[p + "neck.3.weight"]), _p(sd[p + "neck.3.bias"])
        self.model = m
        self.grid = c["image_size"] // c["patch_size"]

    def forward(self, images):
        """images [B,3,S,S] -> bf16 [B, grid*grid, out_chans] (token-major image embeddings)."""
        lib = _lib.load()
        images = images.to(bf16).contiguous()
        B = images.shape[0]
        if B not in self._maps:
            self._maps[B] = _window_maps(B, self.grid, self.WINDOW, images.device)
        part, unpart, nw = self._maps[B]
        ws = self.ws.get(lib.mpl_sam_encoder_workspace_bytes(ctypes.byref(self.model), B), images.device)
        out = torch.empty((B, self.grid * self.grid, self.cfg["out_chans"]), dtype=bf16, device=images.device)
        _lib.check(lib.mpl_sam_encoder_forward(ctypes.byref(self.model), _vp(images), B, _vp(part), _vp(unpart), nw,
                                               _vp(out), _vp(ws), ctypes.c_longlong(ws.numel()), _stream()),
                   "mpl_sam_encoder_forward")
        return out


def dense_pe(gauss, grid):
    """PromptEncoder.get_dense_pe (prompt_encoder.py:62-71,204-226), token-major fp32 [grid*grid, 2F]. Input
    independent, so it is computed once per model on the host side."""
    G = gauss.to(torch.float32)
    ones = torch.ones((grid, grid), dtype=torch.float32, device=G.device)
    y = (ones.cumsum(0) - 0.5) / grid
    x = (ones.cumsum(1) - 0.5) / grid
    c = (2 * torch.stack([x, y], dim=-1) - 1) @ G
    c = 2 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1).reshape(grid * grid, -1).contiguous()


class MaskDecoderEngine:
    """Prompt encoder (text path) + mask decoder (prefix e.g. 'model.visual_model.')."""

    def __init__(self, sd, prefix, grid=16):
        self.prefix, self.grid = prefix, grid
        self.ws = _Workspace()
        self.refresh(sd)

    def refresh(self, sd):
        with _collect(self):
            self._refresh(sd)

    def _refresh(self, sd):
        p = self.prefix + "mask_decoder."
        pe = self.prefix + "prompt_encoder."
        self._keep = []

        def keep(t):
            t = t.contiguous()
            self._keep.append(t)
            return t.data_ptr()

        def attn(a, ap):
            for nm, w, b in (("q_proj", "q_w", "q_b"), ("k_proj", "k_w", "k_b"), ("v_proj", "v_w", "v_b"),
                             ("out_proj", "o_w", "o_b")):
                setattr(a, w, _p(sd[f"{ap}{nm}.weight"]))
                setattr(a, b, _p(sd[f"{ap}{nm}.bias"]))

        D = sd[p + "iou_token.weight"].shape[1]
        depth = 1 + max(int(k[len(p + "transformer.layers."):].split(".")[0]) for k in sd
                        if k.startswith(p + "transformer.layers."))
        layers = (_lib.SamTwoWayLayer * depth)()
        for i in range(depth):
            lp = f"{p}transformer.layers.{i}."
            lay = layers[i]
            attn(lay.self_attn, lp + "self_attn.")
            attn(lay.t2i, lp + "cross_attn_token_to_image.")
            attn(lay.i2t, lp + "cross_attn_image_to_token.")
            for j in (1, 2, 3, 4):
                setattr(lay, f"n{j}_w", _p(sd[f"{lp}norm{j}.weight"]))
                setattr(lay, f"n{j}_b", _p(sd[f"{lp}norm{j}.bias"]))
            lay.lin1_w, lay.lin1_b = _p(sd[lp + "mlp.lin1.weight"]), _p(sd[lp + "mlp.lin1.bias"])
            lay.lin2_w, lay.lin2_b = _p(sd[lp + "mlp.lin2.weight"]), _p(sd[lp + "mlp.lin2.bias"])
        self.layers = layers
        m = _lib.SamMaskDecoder()
        m.dim, m.depth, m.grid = D, depth, self.grid
        m.n_heads = 8
        m.mlp = sd[p + "transformer.layers.0.mlp.lin1.weight"].shape[0]
        m.n_mask_tokens = sd[p + "mask_tokens.weight"].shape[0]
        m.iou_token, m.mask_tokens = _p(sd[p + "iou_token.weight"]), _p(sd[p + "mask_tokens.weight"])
        m.no_mask = _p(sd[pe + "no_mask_embed.weight"])
        m.dense_pe = keep(dense_pe(sd[pe + "pe_layer.positional_encoding_gaussian_matrix"], self.grid))
        m.layers = layers
        attn(m.final_attn, p + "transformer.final_attn_token_to_image.")
        m.nf_w, m.nf_b = _p(sd[p + "transformer.norm_final_attn.weight"]), _p(sd[p + "transformer.norm_final_attn.bias"])
        # ConvTranspose2d(k2,s2) weight [Cin, Cout, 2, 2] -> GEMM weight [(ky,kx,co), ci]; bias tiled over the 4 taps
        w0, b0 = sd[p + "output_upscaling.0.weight"], sd[p + "output_upscaling.0.bias"]
        w1, b1 = sd[p + "output_upscaling.3.weight"], sd[p + "output_upscaling.3.bias"]
        m.up0_w = keep(w0.permute(2, 3, 1, 0).reshape(-1, w0.shape[0]))
        m.up0_b = keep(b0.repeat(4))
        m.up_ln_w, m.up_ln_b = _p(sd[p + "output_upscaling.1.weight"]), _p(sd[p + "output_upscaling.1.bias"])
        m.up1_w = keep(w1.permute(2, 3, 1, 0).reshape(-1, w1.shape[0]))
        m.up1_b = keep(b1.repeat(4))
        g = self.grid
        Y, X = torch.meshgrid(torch.arange(4 * g), torch.arange(4 * g), indexing="ij")
        src = ((((Y // 4) * g + (X // 4)) * 4 + ((Y // 2) % 2) * 2 + (X // 2) % 2) * 4 + (Y % 2) * 2 + (X % 2))
        m.shuffle_idx = keep(src.reshape(-1).to(torch.int32).to(w0.device))
        for j in range(3):
            m.hyper_w[j] = _p(sd[f"{p}output_hypernetworks_mlps.0.layers.{j}.weight"])
            m.hyper_b[j] = _p(sd[f"{p}output_hypernetworks_mlps.0.layers.{j}.bias"])
            m.iou_w[j] = _p(sd[f"{p}iou_prediction_head.layers.{j}.weight"])
            m.iou_b[j] = _p(sd[f"{p}iou_prediction_head.layers.{j}.bias"])
        self.model = m
        self.dim = D

    def forward(self, image_embedding, text_embed):
        """image_embedding bf16 [grid*grid, dim] (token-major), text_embed bf16 [dim] ->
        (low_res_mask bf16 [1,1,4g,4g], iou bf16 [1,1])."""
        lib = _lib.load()
        dev = image_embedding.device
        ws = self.ws.get(lib.mpl_sam_mask_decoder_workspace_bytes(ctypes.byref(self.model)), dev)
        g = self.grid
        mask = torch.empty((1, 1, 4 * g, 4 * g), dtype=bf16, device=dev)
        iou = torch.empty((self.model.n_mask_tokens,), dtype=bf16, device=dev)
        image_embedding = image_embedding.to(bf16).contiguous()
        text_embed = text_embed.to(bf16).contiguous()
        _lib.check(lib.mpl_sam_mask_decoder_forward(ctypes.byref(self.model), _vp(image_embedding), _vp(text_embed),
                                                    _vp(mask), _vp(iou), _vp(ws), ctypes.c_longlong(ws.numel()),
                                                    _stream()), "mpl_sam_mask_decoder_forward")
        return mask, iou[:1].view(1, 1)
