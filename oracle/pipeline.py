"""Oracle: the end-to-end paths of MedPLIBForCausalLM (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Functional restatement of model/MedPLIB.py::evaluate (:574-680) and ::model_forward(inference=True) (:364-511, with
the App. B-12 fix), composed from the pinned pieces: CLIP tower (oracle/clip.py), projector / splice (oracle/arch.py),
LLaMA-MoE stack + greedy search (oracle/llama.py, oracle/moe.py), [SEG] head + postprocess (oracle/heads.py) and
SAM-Med2D (oracle/sam.py). ``sd`` is a state dict with the reference's parameter names.
"""
import torch

from . import arch, clip, heads, llama, sam

VT = "model.vision_tower.vision_tower."
SAM = "model.visual_model."


def encode_images(sd, cfg, images_clip):
    feats = clip.vision_tower(sd, VT, images_clip, cfg["clip"], select_layer=cfg.get("mm_vision_select_layer", -2))
    x = arch.mm_projector(sd, "model.mm_projector.", feats)
    if cfg.get("mm_token_compress"):
        x = arch.token_compressor(sd, "model.mm_token_compressor.", x, cfg.get("mm_compressed_token_count", 256))
    return feats, x


def prefill_inputs(sd, cfg, images_clip, input_ids, attention_mask):
    _, x = encode_images(sd, cfg, images_clip)
    emb, _, am = arch.splice(sd["model.embed_tokens.weight"], input_ids, None, attention_mask, x,
                             use_im_start_end=cfg.get("mm_use_im_start_end", True))
    return emb, am, x.shape[1]


def decode_masks(sd, cfg, pred_embeddings, images, resize_list, size_list):
    emb = sam.image_encoder(sd, SAM + "image_encoder.", images, num_heads=cfg["sam"]["num_heads"])
    g = emb.shape[-1]
    dpe = sam.dense_pe(sd, SAM + "prompt_encoder.", (g, g))
    out, low_all = [], []
    for i in range(len(pred_embeddings)):
        text = pred_embeddings[i].unsqueeze(0).unsqueeze(1)
        sparse, dense = sam.prompt_encoder_text(sd, SAM + "prompt_encoder.", text, (g, g))
        sparse = sparse.to(pred_embeddings.dtype)
        low, _ = sam.mask_decoder(sd, SAM + "mask_decoder.", emb[i].unsqueeze(0), dpe, sparse, dense, False)
        low_all.append(low)
        out.append(heads.postprocess_masks(low, resize_list[i], size_list[i])[:, 0])
    return out, low_all


def evaluate(sd, cfg, images_clip, images, input_ids, resize_list, size_list, max_new_tokens, seg_token_idx,
             attention_mask=None, forced_tokens=None):
    """Returns dict(output_ids, pred_masks, low_res, step_logits, hidden)."""
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids, dtype=torch.bool)
    emb, am, n_img = prefill_inputs(sd, cfg, images_clip, input_ids, attention_mask)
    embed_w = sd["model.embed_tokens.weight"]
    new, hidden, step_logits = llama.greedy_generate(sd, cfg["llama"], emb, am, lambda ids: embed_w[ids],
                                                     max_new_tokens, forced_tokens=forced_tokens)
    output_ids = torch.cat([input_ids, new], dim=1)
    mask = heads.seg_token_mask(output_ids, seg_token_idx, n_img)[:, :hidden.shape[1]]
    rows = hidden[mask]
    if rows.shape[0] > 1:
        rows = rows[:1]
    elif rows.shape[0] == 0:
        rows = hidden[:1, -2:-1, :].squeeze(1)
    pred = heads.text_hidden_fcs(sd, "model.text_hidden_fcs.0.", rows)
    masks, low = decode_masks(sd, cfg, pred, images, resize_list, size_list)
    return dict(output_ids=output_ids, pred_masks=masks, low_res=low, step_logits=step_logits, hidden=hidden,
                pred_embeddings=pred)


def grounding_forward(sd, cfg, images_clip, images, input_ids, resize_list, size_list, seg_token_idx,
                      attention_mask=None):
    """model_forward(inference=True): single pass with <SEG> in the prompt."""
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids, dtype=torch.bool)
    emb, am, n_img = prefill_inputs(sd, cfg, images_clip, input_ids, attention_mask)
    out = llama.model_forward(sd, cfg["llama"], emb, am)
    hidden = out["last_hidden_state"]
    mask = heads.seg_token_mask(input_ids, seg_token_idx, n_img)[:, :hidden.shape[1]]
    pred = heads.text_hidden_fcs(sd, "model.text_hidden_fcs.0.", hidden[mask])
    masks, low = decode_masks(sd, cfg, pred, images, resize_list, size_list)
    return dict(pred_masks=masks, low_res=low, hidden=hidden, pred_embeddings=pred)


def grounding_forward_icl(sd, cfg, images_clip_list, mask_images_list, image_token_types, image_token_lengths, images,
                          input_ids, resize_list, size_list, seg_token_idx, attention_mask=None):
    """model_forward(inference=True) in MedPLIB-ICL separate mode (medplib_arch.py:246-266): every sample brings a
    stack of CLIP images (exemplars + query) and a stack of exemplar masks; each IMAGE sentinel of the prompt is
    replaced, in order, by the compressed image tokens (mm_token_compress) or the MaskTokenEncoder tokens named by
    image_token_types; the [SEG] mask skips image_token_lengths entries per sentinel (MedPLIB.py:310-355)."""
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids, dtype=torch.bool)
    _, x = encode_images(sd, cfg, torch.cat(list(images_clip_list), dim=0))
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
    emb, _, am = arch.splice(sd["model.embed_tokens.weight"], input_ids, None, attention_mask, combined,
                             use_im_start_end=cfg.get("mm_use_im_start_end", True), per_token_features=True)
    out = llama.model_forward(sd, cfg["llama"], emb, am)
    hidden = out["last_hidden_state"]
    mask = heads.seg_token_mask(input_ids, seg_token_idx, x.shape[1], image_token_lengths)[:, :hidden.shape[1]]
    pred = heads.text_hidden_fcs(sd, "model.text_hidden_fcs.0.", hidden[mask])
    masks, low = decode_masks(sd, cfg, pred, images, resize_list, size_list)
    return dict(pred_masks=masks, low_res=low, hidden=hidden, pred_embeddings=pred, inputs_embeds=emb)
