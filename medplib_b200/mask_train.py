"""Grounding head of the train step (seg_flag with inference=False): text_hidden_fcs on the [SEG] rows -> SAM-Med2D
prompt encoder (text) + two-way mask decoder -> postprocess_masks -> BCE / Dice / IoU / Focal, differentiable.
Replaces torch autograd over model/MedPLIB.py:456-559 and model/segment_anything_med2d/modeling/
{mask_decoder.py:71-153, transformer.py:16-244, prompt_encoder.py:140-187}.

Every arithmetic step is one of our kernels wrapped in a torch.autograd.Function (forward = the inference kernels of
ops.py, backward = medplib_b200/csrc/mask_train.cu); torch.autograd is only the tape: it orders the backward calls and
sums fan-in gradients of the <= 7x256 token tensors. Weight gradients never become ``.grad`` tensors: each Function adds
them straight into the Trainer's fp32 arena.
"""
import math

import torch

from . import engine, ops
from . import train_ops as T

bf16 = torch.bfloat16
f32 = torch.float32


def _acc(tr, param, grad):
    """arena[param] += grad (any float dtype, same number of elements)."""
    g = tr.arena.of(param)
    if g is not None and grad is not None:
        grad = grad.contiguous()
        if grad.numel() == g.numel():
            grad = grad.reshape(1, -1)
        T.col_sum(grad, g.view(-1))  # 2-D [m, numel]: the m rows are summed (tiled conv bias)


class ParamFn(torch.autograd.Function):
    """A trainable tensor entering the tape by value (tokens, repacked conv weights): backward adds into the arena."""

    @staticmethod
    def forward(ctx, param, tr, transform, inverse):
        ctx.tr, ctx.param, ctx.inverse = tr, param, inverse
        v = param.detach()
        return transform(v) if transform is not None else v.clone()

    @staticmethod
    def backward(ctx, dy):
        g = dy.contiguous()
        if ctx.inverse is not None:
            g = ctx.inverse(g)
        _acc(ctx.tr, ctx.param, g)
        return None, None, None, None


def param(tr, p, transform=None, inverse=None):
    if not p.requires_grad:
        v = p.detach()
        return transform(v) if transform is not None else v
    return ParamFn.apply(p, tr, transform, inverse)


class LinFn(torch.autograd.Function):
    """y = act(x W^T + b), act in {None, relu}. W: an nn.Parameter (gradient -> arena) or a tape tensor (gradient
    returned). x [M, K] bf16."""

    @staticmethod
    def forward(ctx, x, W, b, act, tr):
        x = x.contiguous()
        Wd = W.detach().contiguous()
        y = ops.linear(x, Wd, bias=b.detach() if b is not None else None, act=act)
        ctx.tr, ctx.act, ctx.W, ctx.b = tr, act, W, b
        ctx.save_for_backward(x, Wd, y if act == "relu" else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Wd, y = ctx.saved_tensors
        tr = ctx.tr
        dy = dy.contiguous()
        if ctx.act == "relu":
            dy = T.act_bwd(y, dy, "relu")
        dx = T.gemm_small(dy, Wd) if ctx.needs_input_grad[0] else None
        dW = None
        gW = tr.arena.of(ctx.W)
        if gW is not None:
            T.gemm_small(dy, x, out=gW, trans_a=True, accumulate=True)
        elif ctx.needs_input_grad[1]:  # tape tensor: fp32 accumulate (split-K capable), handed back in bf16
            dW32 = torch.zeros((dy.shape[1], x.shape[1]), dtype=f32, device=dy.device)
            dW = T.gemm_small(dy, x, out=dW32, trans_a=True, accumulate=True).to(bf16)
        if ctx.b is not None:
            gb = tr.arena.of(ctx.b)
            if gb is not None:
                T.col_sum(dy, gb)
            elif ctx.needs_input_grad[2]:
                db = torch.zeros(dy.shape[1], dtype=f32, device=dy.device)
                T.col_sum(dy, db)
                return dx, dW, db.to(bf16), None, None
        return dx, dW, None, None, None


def lin(tr, x, W, b=None, act=None):
    return LinFn.apply(x, W, b, act, tr)


def lin_mod(tr, x, mod, act=None):
    return LinFn.apply(x, mod.weight, mod.bias, act, tr)


class LnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps, tr):
        x = x.contiguous()
        ctx.tr, ctx.w, ctx.b, ctx.eps = tr, w, b, eps
        ctx.save_for_backward(x)
        return ops.layernorm(x, w.detach(), b.detach(), eps)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        ar = ctx.tr.arena
        dx = T.layernorm_bwd(x, ctx.w.detach(), dy.contiguous(), ctx.eps, dweight=ar.of(ctx.w), dbias=ar.of(ctx.b))
        return dx, None, None, None, None


def ln(tr, x, mod, eps=None):
    return LnFn.apply(x, mod.weight, mod.bias, mod.eps if eps is None else eps, tr)


class AttnFn(torch.autograd.Function):
    """softmax(q k^T / sqrt(d)) v per (sample, head); q [B*Tq, H*d], k / v [B*Tk, H*d] (samples stacked along rows)."""

    @staticmethod
    def forward(ctx, q, k, v, H, B):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        C = q.shape[1]
        Tq, Tk = q.shape[0] // B, k.shape[0] // B
        d = C // H
        ctx.H, ctx.B, ctx.scale = H, B, 1.0 / math.sqrt(d)
        o = ops.attention(q.view(B, Tq, H, d), k.view(B, Tk, H, d), v.view(B, Tk, H, d), ctx.scale)
        ctx.save_for_backward(q, k, v)
        return o.view(B * Tq, C)

    @staticmethod
    def backward(ctx, do):
        q, k, v = ctx.saved_tensors
        dq, dk, dv = T.attn_small_bwd(q, k, v, do.contiguous(), ctx.H, ctx.scale, batch=ctx.B)
        return dq, dk, dv, None, None


class AddFn(torch.autograd.Function):
    """bf16(a + b); b is a same-shape tape tensor or a constant (bf16 / f32, broadcast over rows)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.same = b.shape == a.shape
        return ops.add(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return dy, (dy if ctx.same and ctx.needs_input_grad[1] else None)


def add(a, b):
    return AddFn.apply(a, b)


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return T.act_fwd(x, "gelu")

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return T.act_bwd(x, dy.contiguous(), "gelu")


class RowsFn(torch.autograd.Function):
    """out[r] = x[idx[r]] (0 where idx == -1); backward gathers with the inverse map."""

    @staticmethod
    def forward(ctx, x, idx, inv):
        ctx.inv, ctx.shape = inv, x.shape
        return ops.gather_rows(idx, table=x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return ops.gather_rows(ctx.inv, table=dy.contiguous()).view(ctx.shape), None, None


class BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size):
        ctx.in_hw = tuple(x.shape[-2:])
        return ops.bilinear_resize(x, tuple(size))

    @staticmethod
    def backward(ctx, dy):
        return T.bilinear_resize_bwd(dy, ctx.in_hw), None


class MaskLossFn(torch.autograd.Function):
    """-> f32 [4] = (sigmoid_ce_loss, dice_loss, MaskIoULoss, FocalLoss) of ONE mask (MedPLIB.py:26-124)."""

    @staticmethod
    def forward(ctx, pred, pred_iou, gt):
        pred, pred_iou = pred.contiguous(), pred_iou.contiguous()
        out, sums = T.mask_losses(pred, gt, pred_iou)
        ctx.save_for_backward(pred, pred_iou, gt, sums)
        return out

    @staticmethod
    def backward(ctx, d4):
        pred, pred_iou, gt, sums = ctx.saved_tensors
        dpred, dpi = T.mask_losses_bwd(pred, gt, pred_iou, sums, d4.float())
        return dpred, dpi.to(bf16).view(pred_iou.shape), None


# ----------------------------------------------------------------------------------------------------------------
def select_rows(hidden, seg_mask):
    """hidden [B,T,D] (tape) -> rows where seg_mask (bool [B,T]) is set, in row-major order (`hidden[seg_mask]`)."""
    B, Tn, D = hidden.shape
    flat = seg_mask.reshape(-1)
    pos = flat.nonzero().flatten()  # one small host sync (the reference's boolean indexing does the same)
    idx = pos.to(torch.int32)
    inv = torch.full((B * Tn,), -1, dtype=torch.int32, device=hidden.device)
    inv[pos] = torch.arange(pos.numel(), dtype=torch.int32, device=hidden.device)
    return RowsFn.apply(hidden.reshape(B * Tn, D), idx, inv)


def text_hidden_fcs(tr, fc, rows):
    """Sequential(Linear, ReLU, Linear, Dropout(0)) of MedPLIB.py:153-164 on the [SEG] rows."""
    return lin_mod(tr, lin_mod(tr, rows, fc[0], act="relu"), fc[2])


def _attention(tr, a, q_in, k_in, v_in, H, B):
    q = lin_mod(tr, q_in, a.q_proj)
    k = lin_mod(tr, k_in, a.k_proj)
    v = lin_mod(tr, v_in, a.v_proj)
    return lin_mod(tr, AttnFn.apply(q, k, v, H, B), a.out_proj)


def _consts(tr, model, B):
    c = getattr(tr, "_mask_consts", None)
    if c is None:
        vm = model.model.visual_model
        grid = vm.prompt_encoder.image_embedding_size[0]
        dev = model.lm_head.weight.device
        gauss = vm.prompt_encoder.pe_layer.positional_encoding_gaussian_matrix
        Y, X = torch.meshgrid(torch.arange(4 * grid), torch.arange(4 * grid), indexing="ij")
        src = ((((Y // 4) * grid + (X // 4)) * 4 + ((Y // 2) % 2) * 2 + (X // 2) % 2) * 4 + (Y % 2) * 2 + (X % 2))
        shuffle = src.reshape(-1)
        inv = torch.empty_like(shuffle)
        inv[shuffle] = torch.arange(shuffle.numel())
        c = dict(grid=grid, dense_pe=engine.dense_pe(gauss.detach(), grid).to(dev), shuffle=shuffle, shuffle_inv=inv,
                 dev=dev, per_batch={})
        tr._mask_consts = c
    if B not in c["per_batch"]:
        P = c["shuffle"].numel()
        off = (torch.arange(B) * P).repeat_interleave(P)
        c["per_batch"][B] = dict(
            dense_pe=c["dense_pe"].repeat(B, 1).contiguous(),
            shuffle=(c["shuffle"].repeat(B) + off).to(torch.int32).to(c["dev"]),
            shuffle_inv=(c["shuffle_inv"].repeat(B) + off).to(torch.int32).to(c["dev"]))
    return c, c["per_batch"][B]


def mask_decoder(tr, model, img_tok, text):
    """img_tok bf16 [B, g*g, D] (frozen image embeddings, token-major), text [B, D] (tape) -> (low_res list of
    [1, 4g, 4g], iou list of [1]); multimask_output=False (mask / IoU of token 0), like MedPLIB.py:488-495. The B masks
    of a step run as ONE batch (samples stacked along the row dimension of every kernel)."""
    vm = model.model.visual_model
    md, tf = vm.mask_decoder, vm.mask_decoder.transformer
    B, Tn, D = img_tok.shape
    c, cb = _consts(tr, model, B)
    g = c["grid"]
    H = 8
    base = torch.cat([param(tr, md.iou_token.weight), param(tr, md.mask_tokens.weight)], dim=0)  # [5, D]
    nt = base.shape[0] + 1
    tokens0 = torch.cat([base.unsqueeze(0).expand(B, -1, -1), text.view(B, 1, D)], dim=1).reshape(B * nt, D)
    keys = ops.add(img_tok.reshape(B * Tn, D).contiguous(), vm.prompt_encoder.no_mask_embed.weight.detach().reshape(-1))
    dpe = cb["dense_pe"]
    tok, tpe = tokens0, tokens0

    def token_to_image(tok, keys, a, norm):
        o = _attention(tr, a, add(tok, tpe), add(keys, dpe), keys, H, B)
        return ln(tr, add(tok, o), norm)

    for i, L in enumerate(tf.layers):
        if i == 0:
            tok = _attention(tr, L.self_attn, tok, tok, tok, H, B)
        else:
            qk = add(tok, tpe)
            tok = add(tok, _attention(tr, L.self_attn, qk, qk, tok, H, B))
        tok = ln(tr, tok, L.norm1)
        tok = token_to_image(tok, keys, L.cross_attn_token_to_image, L.norm2)
        m = lin_mod(tr, lin_mod(tr, tok, L.mlp.lin1, act="relu"), L.mlp.lin2)
        tok = ln(tr, add(tok, m), L.norm3)
        o = _attention(tr, L.cross_attn_image_to_token, add(keys, dpe), add(tok, tpe), tok, H, B)
        keys = ln(tr, add(keys, o), L.norm4)
    tok = token_to_image(tok, keys, tf.final_attn_token_to_image, tf.norm_final_attn)

    # upscaling: ConvTranspose2d(k2,s2) = one GEMM against the [(ky,kx,co), ci] repack of its weight, rows become
    # (sample, pixel, tap); LayerNorm2d = LayerNorm over the channel rows; final row gather into raster order
    up = md.output_upscaling
    C4, C8 = D // 4, D // 8

    def convt_w(conv):
        ci, co = conv.weight.shape[:2]
        w = param(tr, conv.weight, lambda v: v.permute(2, 3, 1, 0).reshape(4 * co, ci).contiguous(),
                  lambda gr: gr.view(2, 2, co, ci).permute(3, 2, 0, 1).contiguous())
        b = param(tr, conv.bias, lambda v: v.repeat(4), lambda gr: gr.view(4, co))
        return w, b

    w0, b0 = convt_w(up[0])
    w1, b1 = convt_w(up[3])
    u0 = lin(tr, keys, w0, b0).view(4 * B * Tn, C4)
    u0 = GeluFn.apply(ln(tr, u0, up[1], eps=1e-6))
    u1 = GeluFn.apply(lin(tr, u0, w1, b1)).view(16 * B * Tn, C8)
    ups = RowsFn.apply(u1, cb["shuffle"], cb["shuffle_inv"])  # [B * (4g)^2, C8], raster order per sample
    P = 16 * Tn

    tok3 = tok.view(B, nt, D)
    hy = md.output_hypernetworks_mlps[0].layers
    hv = lin_mod(tr, lin_mod(tr, lin_mod(tr, tok3[:, 1], hy[0], act="relu"), hy[1], act="relu"), hy[2])  # [B, C8]
    io = md.iou_prediction_head.layers
    iou = lin_mod(tr, lin_mod(tr, lin_mod(tr, tok3[:, 0], io[0], act="relu"), io[1], act="relu"), io[2])  # [B, n_mask]
    lows = [lin(tr, hv[b:b + 1], ups[b * P:(b + 1) * P]).view(1, 4 * g, 4 * g) for b in range(B)]  # hyper_in @ upscaled
    return lows, [iou[b, 0:1] for b in range(B)]


def mask_head_losses(tr, model, pred_embeddings, image_embeddings, resize_list, size_list, masks_list):
    """MedPLIB.py:473-545: one mask per [SEG] embedding; returns the SUMS over masks of the four losses (tape f32
    scalars) and num_masks; the caller divides and weights like :547-559."""
    n = len(pred_embeddings)
    Bi, C, g, _ = image_embeddings.shape
    tok = image_embeddings.permute(0, 2, 3, 1).reshape(Bi, g * g, C)
    sums = {"bce": 0, "dice": 0, "iou": 0, "focal": 0, "num_masks": 0}
    pred_masks = []
    if n == 0:
        sums["pred_masks"] = pred_masks
        return sums
    lows, ious = mask_decoder(tr, model, tok[:n], pred_embeddings)
    for i in range(n):
        low, iou = lows[i], ious[i]
        inp, orig = resize_list[i], size_list[i]
        pad_h, pad_w = low.shape[-2] - inp[0], low.shape[-1] - inp[1]
        top, left = pad_h // 2, pad_w // 2
        oh, ow = low.shape[-2] - pad_h, low.shape[-1] - pad_w
        crop = low[:, top:top + oh, left:left + ow]
        if crop.shape != low.shape:
            crop = crop.contiguous()
        pm = BilinearFn.apply(crop, tuple(orig))  # [1, H, W]
        pred_masks.append(pm)
        gt = masks_list[i].to(f32).reshape(1, *orig).contiguous()
        assert gt.shape[0] == pm.shape[0], f"gt_mask.shape: {gt.shape}, pred_mask.shape: {pm.shape}"
        l4 = MaskLossFn.apply(pm, iou, gt)
        k = gt.shape[0]
        sums["bce"] = sums["bce"] + l4[0] * k
        sums["dice"] = sums["dice"] + l4[1] * k
        sums["iou"] = sums["iou"] + l4[2] * k
        sums["focal"] = sums["focal"] + l4[3] * k
        sums["num_masks"] += k
    sums["pred_masks"] = pred_masks
    return sums
