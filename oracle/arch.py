"""Oracle: multimodal glue (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Functional restatement of model/medplib/model/medplib_arch.py: mm_projector (multimodal_projector/builder.py:39-46,
``mlp2x_gelu``), TokenCompressor :67-77, MaskTokenEncoder :80-108, region_fea_adapter + extract_region_feature
:580-614 + point_sample :39-64, encode_images :198-212 and the sentinel splice prepare_inputs_labels_for_multimodal
:217-527 (the non-``tune_mm_mlp_adapter`` branches, which are the ones MedPLIB configures: MedPLIB.py:176-184).
Pinned against the reference's own classes through tests/golden/arch_*.pt.
"""
import math

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
REGION_TOKEN_INDEX = -300


def mm_projector(sd, p, x):
    h = F.gelu(F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"]))
    return F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"])


def token_compressor(sd, p, x, num_tokens):
    x = F.adaptive_avg_pool1d(x.transpose(1, 2), num_tokens).transpose(1, 2)
    x = F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def mask_token_encoder(sd, p, masks, num_tokens):
    if masks.dim() == 3:
        masks = masks.unsqueeze(1)
    masks = masks[:, :1].to(sd[p + "proj.weight"].dtype)
    x = masks
    for i in (0, 2, 4, 6):
        x = F.gelu(F.conv2d(x, sd[f"{p}encoder.{i}.weight"], sd[f"{p}encoder.{i}.bias"], stride=2, padding=1))
    x = F.adaptive_avg_pool1d(x.flatten(2), num_tokens).transpose(1, 2)
    x = F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])
    return F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)


def region_features(region_feature_map, region_masks, max_sample_point, original_dtype, return_dtype):
    """extract_region_feature: per image, per region mask -> mean of bilinearly sampled (align_corners=True, fp32)
    adapter features at the mask's non-zero pixels. Regions with more than max_sample_point pixels use randperm in the
    reference (not reproducible) — the oracle requires <= max_sample_point."""
    out = []
    for fmap, masks in zip(region_feature_map, region_masks):
        if len(masks) == 0:
            out.append(None)
            continue
        hw = torch.tensor([masks[0].shape[0], masks[0].shape[1]])[None]
        pts = []
        for m in masks:
            nz = m.nonzero() / hw
            assert nz.shape[0] <= max_sample_point, "oracle: region larger than max_sample_point (randperm path)"
            pts.append(nz)
        pos = torch.nn.utils.rnn.pad_sequence(pts, padding_value=-1, batch_first=True)
        valid = ~(pos.sum(dim=-1) < 0)
        h = w = int(math.sqrt(fmap.shape[0]))
        c = fmap.shape[-1]
        f = fmap.reshape(h, w, c).permute(2, 0, 1).unsqueeze(0).repeat(pos.shape[0], 1, 1, 1).to(original_dtype)
        coords = pos.flip(dims=(2,)).type(original_dtype).unsqueeze(2)
        s = F.grid_sample(f.float(), (2.0 * coords - 1.0).float(), align_corners=True).to(return_dtype).squeeze(3)
        s = s.to(fmap.dtype)
        out.append(torch.stack([x[m].mean(dim=0) for x, m in zip(s.transpose(1, 2), valid)]).nan_to_num())
    return out


def splice(embed_w, input_ids, labels, attention_mask, image_features, region_feats=None, valid_region=None,
           use_im_start_end=True, per_token_features=False):
    """prepare_inputs_labels_for_multimodal :296-527.

    image_features: list (one entry per IMAGE sentinel, in order) when per_token_features, else tensor [B, n, D] (one
    image per sample). region_feats: list per VALID sample of [n_regions, D]; valid_region: list of bool per sample.
    Returns (inputs_embeds [B,T,D], labels [B,T] or None, attention_mask [B,T] or None).
    """
    new_embeds, new_labels = [], ([] if labels is not None else None)
    img_idx = 0
    for b, ids in enumerate(input_ids):
        if (ids == IMAGE_TOKEN_INDEX).sum() == 0:
            new_embeds.append(embed_w[ids])
            if labels is not None:
                new_labels.append(labels[b])
            if not per_token_features:
                img_idx += 1
            continue
        pieces, lab_pieces = [], []
        cur_labels = labels[b] if labels is not None else None
        pos = torch.where(ids == IMAGE_TOKEN_INDEX)[0]
        while pos.numel() > 0:
            feats = image_features[img_idx]
            s = int(pos[0])
            pieces.append(embed_w[ids[:s]])
            pieces.append(feats)
            if use_im_start_end:
                pieces.append(embed_w[ids[s + 1:s + 2]])
            if labels is not None:
                lab_pieces.append(cur_labels[:s])
                lab_pieces.append(torch.full((feats.shape[0],), IGNORE_INDEX, dtype=labels.dtype))
                if use_im_start_end:
                    lab_pieces.append(cur_labels[s + 1:s + 2])
                    cur_labels = cur_labels[s + 2:]
                else:
                    cur_labels = cur_labels[s + 1:]
            img_idx += 1
            ids = ids[s + 2:] if use_im_start_end else ids[s + 1:]
            pos = torch.where(ids == IMAGE_TOKEN_INDEX)[0]
        if ids.numel() > 0:
            ridx = (ids == REGION_TOKEN_INDEX).nonzero(as_tuple=True)[0].tolist()
            ids = ids[ids != REGION_TOKEN_INDEX]
            text = embed_w[ids]
            if labels is not None:
                lab_pieces.append(cur_labels)
            if region_feats is not None and valid_region is not None and valid_region[b]:
                k = sum(bool(v) for v in valid_region[:b + 1]) - 1
                for j, at in enumerate(ridx):
                    text = torch.cat((text[:at], region_feats[k][j].unsqueeze(0), text[at:]))
            pieces.append(text)
        new_embeds.append(torch.cat(pieces, dim=0))
        if labels is not None:
            new_labels.append(torch.cat(lab_pieces, dim=0))
    max_len = max(x.shape[0] for x in new_embeds)
    ragged = any(x.shape[0] != max_len for x in new_embeds)
    D = new_embeds[0].shape[1]
    emb = torch.stack([torch.cat((x, torch.zeros((max_len - x.shape[0], D), dtype=x.dtype))) for x in new_embeds])
    lab = None
    if labels is not None:
        lab = torch.stack([torch.cat((x, torch.full((max_len - x.shape[0],), IGNORE_INDEX, dtype=x.dtype)))
                           for x in new_labels])
    am = None
    if attention_mask is not None:
        rows = []
        for b in range(len(new_embeds)):
            n_b = new_embeds[b].shape[0] if ragged else max_len
            left = torch.ones((n_b - input_ids.shape[1],), dtype=attention_mask.dtype) if ragged else \
                torch.ones((max_len - input_ids.shape[1],), dtype=attention_mask.dtype)
            right = torch.zeros((max_len - n_b,), dtype=attention_mask.dtype)
            rows.append(torch.cat((left, attention_mask[b], right)))
        am = torch.stack(rows)
    return emb, lab, am
