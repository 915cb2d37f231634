"""Oracle: the training forward of MedPLIBForCausalLM.model_forward(inference=False) (TEST INFRASTRUCTURE ONLY — see
oracle/__init__.py). Plain differentiable PyTorch on CPU: torch.autograd over this function is the gradient reference
for the hand-written backward kernels (medplib_b200/train.py, mask_train.py).

Follows model/MedPLIB.py:364-572 (loss assembly :515-572), medplib_moe_llama.py:381-421 (shifted CE + router aux loss),
peft 0.10 LoRA Linear (oracle/llama.py::linear), DeepSpeed top-1 gating in training mode (capacity_factor, RTS
uniforms injected). The frozen encoders (CLIP tower, SAM-Med2D image encoder) run under no_grad like the reference
(clip_encoder.py:41, MedPLIB.py:274-285).
"""
import torch

from . import arch, heads, llama, pipeline, sam


def train_losses(sd, cfg, images_clip, images, input_ids, labels, attention_mask, masks_list, label_sizes,
                 resize_list, seg_token_idx, w, seg_flag=True, rts_uniforms=None, region_masks=None,
                 valid_region=None):
    """Returns (10-key loss dict, aux dict with routing / hidden states for diagnostics). ``w`` = dict(ce, bce, dice,
    iou, focal) loss weights (MedPLIB.py:233-240)."""
    with torch.no_grad():
        feats, _ = pipeline.encode_images(sd, cfg, images_clip)
    # the projector is differentiable (scripts/train_stage2.sh trains it); the CLIP tower is frozen
    x = arch.mm_projector(sd, "model.mm_projector.", feats)
    if cfg.get("mm_token_compress"):
        x = arch.token_compressor(sd, "model.mm_token_compressor.", x, cfg.get("mm_compressed_token_count", 256))
    with torch.no_grad():
        if seg_flag:
            image_emb = sam.image_encoder(sd, pipeline.SAM + "image_encoder.", images,
                                          num_heads=cfg["sam"]["num_heads"])
    region_feats = None
    if region_masks is not None and len(region_masks) > 0:
        # medplib_arch.py:208,426-429,580-614: region_fea_adapter on the raw CLIP features of the valid samples, then
        # point-sampled means spliced at the -300 slots (differentiable w.r.t. the adapter)
        import torch.nn.functional as F
        rmap = F.linear(feats, sd["model.region_fea_adapter.weight"], sd["model.region_fea_adapter.bias"])
        rmap = rmap[torch.tensor([bool(v) for v in valid_region])]
        region_feats = arch.region_features(rmap, region_masks, 512, rmap.dtype, rmap.dtype)
    emb, lab, am = arch.splice(sd["model.embed_tokens.weight"], input_ids, labels, attention_mask, x,
                               region_feats=region_feats, valid_region=valid_region,
                               use_im_start_end=cfg.get("mm_use_im_start_end", True))
    out = llama.model_forward(sd, cfg["llama"], emb, am, training=True, rts_uniforms=rts_uniforms)
    hidden = out["last_hidden_state"]
    logits, loss, moe_loss = llama.causal_lm_tail(sd, cfg["llama"], hidden, lab, out["moe_losses"])
    ce_loss = loss * w["ce"]
    aux = dict(hidden=hidden, gate_logits=out["gate_logits"], logits=logits, labels=lab, moe_loss=moe_loss)
    if not seg_flag:
        z = torch.zeros_like(ce_loss)
        return {"loss": ce_loss, "ce_loss": ce_loss, "mask_bce_loss": z, "mask_dice_loss": z, "mask_loss": z,
                "unscale_mask_bce_loss": z, "unscale_mask_dice_loss": z, "unscale_mask_loss": z,
                "unscale_mask_iou_loss": z, "unscale_mask_focal_loss": z}, aux
    mask = heads.seg_token_mask(input_ids, seg_token_idx, x.shape[1])[:, :hidden.shape[1]]
    pred = heads.text_hidden_fcs(sd, "model.text_hidden_fcs.0.", hidden[mask])
    g = image_emb.shape[-1]
    dpe = sam.dense_pe(sd, pipeline.SAM + "prompt_encoder.", (g, g))
    pred_masks, pred_ious, lows = [], [], []
    for i in range(len(pred)):
        text = pred[i].unsqueeze(0).unsqueeze(1)
        sparse, dense = sam.prompt_encoder_text(sd, pipeline.SAM + "prompt_encoder.", text, (g, g))
        sparse = sparse.to(pred.dtype)
        low, iou = sam.mask_decoder(sd, pipeline.SAM + "mask_decoder.", image_emb[i].unsqueeze(0), dpe, sparse, dense,
                                    False)
        lows.append(low)
        pred_masks.append(heads.postprocess_masks(low, resize_list[i], label_sizes[i])[:, 0])
        pred_ious.append(iou[:, 0])
    aux.update(pred_masks=pred_masks, pred_ious=pred_ious, low_res=lows, pred_embeddings=pred)
    return heads.mask_losses(pred_masks, masks_list, pred_ious, ce_loss, w), aux


def train_losses_icl(sd, cfg, images_clip_list, mask_images_list, image_token_types, image_token_lengths, images,
                     input_ids, labels, attention_mask, masks_list, label_sizes, resize_list, seg_token_idx, w,
                     rts_uniforms=None):
    """model_forward(inference=False) in MedPLIB-ICL separate mode (scripts/train_medplib_icl.sh with
    ICL_MASK_MODE=separate): every sample brings a stack of CLIP images (exemplars + query) and of exemplar masks; the
    IMAGE sentinels are replaced in order by compressed image tokens (mm_token_compressor, trainable) or
    MaskTokenEncoder tokens (mask_encoder, trainable) -- medplib_arch.py:246-266; the [SEG] mask skips
    image_token_lengths entries per sentinel (MedPLIB.py:310-355). Differentiable w.r.t. both adapters (and the
    projector); the CLIP tower and the SAM-Med2D image encoder are frozen."""
    with torch.no_grad():
        feats, _ = pipeline.encode_images(sd, cfg, torch.cat(list(images_clip_list), dim=0))
        image_emb = sam.image_encoder(sd, pipeline.SAM + "image_encoder.", images, num_heads=cfg["sam"]["num_heads"])
    x = arch.mm_projector(sd, "model.mm_projector.", feats)
    x = arch.token_compressor(sd, "model.mm_token_compressor.", x, cfg.get("mm_compressed_token_count", 256))
    mf = arch.mask_token_encoder(sd, "model.mask_encoder.", torch.cat(list(mask_images_list), dim=0),
                                 cfg.get("mask_encoder_token_count", 64))
    combined, ii, mi = [], 0, 0
    for types_ in image_token_types:
        for t in types_:
            if t == "mask":
                combined.append(mf[mi])
                mi += 1
            else:
                combined.append(x[ii])
                ii += 1
    emb, lab, am = arch.splice(sd["model.embed_tokens.weight"], input_ids, labels, attention_mask, combined,
                               use_im_start_end=cfg.get("mm_use_im_start_end", True), per_token_features=True)
    out = llama.model_forward(sd, cfg["llama"], emb, am, training=True, rts_uniforms=rts_uniforms)
    hidden = out["last_hidden_state"]
    logits, loss, moe_loss = llama.causal_lm_tail(sd, cfg["llama"], hidden, lab, out["moe_losses"])
    ce_loss = loss * w["ce"]
    aux = dict(hidden=hidden, gate_logits=out["gate_logits"], logits=logits, labels=lab, moe_loss=moe_loss,
               inputs_embeds=emb)
    mask = heads.seg_token_mask(input_ids, seg_token_idx, x.shape[1], image_token_lengths)[:, :hidden.shape[1]]
    pred = heads.text_hidden_fcs(sd, "model.text_hidden_fcs.0.", hidden[mask])
    g = image_emb.shape[-1]
    dpe = sam.dense_pe(sd, pipeline.SAM + "prompt_encoder.", (g, g))
    pred_masks, pred_ious = [], []
    for i in range(len(pred)):
        text = pred[i].unsqueeze(0).unsqueeze(1)
        sparse, dense = sam.prompt_encoder_text(sd, pipeline.SAM + "prompt_encoder.", text, (g, g))
        low, iou = sam.mask_decoder(sd, pipeline.SAM + "mask_decoder.", image_emb[i].unsqueeze(0), dpe,
                                    sparse.to(pred.dtype), dense, False)
        pred_masks.append(heads.postprocess_masks(low, resize_list[i], label_sizes[i])[:, 0])
        pred_ious.append(iou[:, 0])
    aux.update(pred_masks=pred_masks, pred_ious=pred_ious, pred_embeddings=pred)
    return heads.mask_losses(pred_masks, masks_list, pred_ious, ce_loss, w), aux
