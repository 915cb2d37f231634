"""Oracle: [SEG] head, mask post-processing and the mask losses (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Restates model/MedPLIB.py: text_hidden_fcs :153-164 (applied :456, :631-633), build_seg_token_mask :310-355,
postprocess_masks :682-701, MaskIoULoss :26-47, FocalLoss :49-74, dice_loss :76-109, sigmoid_ce_loss :112-124 and
the loss assembly :515-572. Pinned against the reference's own functions through tests/golden/heads_*.pt.
"""
import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200


def text_hidden_fcs(sd, p, x):
    """Sequential(Linear, ReLU, Linear, Dropout(0)); p = 'model.text_hidden_fcs.0.'"""
    return F.linear(F.relu(F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"])), sd[p + "2.weight"], sd[p + "2.bias"])


def seg_token_mask(input_ids, seg_token_idx, image_token_len, image_token_lengths=None):
    """build_seg_token_mask: position t is marked when token t+1 is <SEG>; each IMAGE sentinel expands to
    image_token_len (or its per-image length) False entries; rows right-padded with False."""
    shifted = torch.zeros_like(input_ids, dtype=torch.bool)
    shifted[:, :-1] = input_ids[:, 1:] == seg_token_idx
    rows = []
    for b in range(input_ids.shape[0]):
        cur, k = [], 0
        for t in range(input_ids.shape[1]):
            if int(input_ids[b, t]) == IMAGE_TOKEN_INDEX:
                n = image_token_len
                if image_token_lengths is not None and len(image_token_lengths) > b and \
                        len(image_token_lengths[b]) > k:
                    n = image_token_lengths[b][k]
                k += 1
                cur.append(torch.zeros(n, dtype=torch.bool))
            else:
                cur.append(shifted[b, t].view(1))
        rows.append(torch.cat(cur))
    T = max(r.shape[0] for r in rows)
    return torch.stack([torch.cat([r, torch.zeros(T - r.shape[0], dtype=torch.bool)]) for r in rows])


def postprocess_masks(masks, input_size, original_size):
    """postprocess_masks: centre-crop by (mask_size - input_size) — negative for any side > mask size, i.e. a no-op at
    64x64 — then bilinear resize (align_corners=False) to the ground-truth size."""
    if masks.dim() == 3:
        masks = masks.unsqueeze(0)
    pad_h = masks.shape[-2] - input_size[0]
    pad_w = masks.shape[-1] - input_size[1]
    top, left = pad_h // 2, pad_w // 2
    oh, ow = masks.shape[-2] - pad_h, masks.shape[-1] - pad_w
    masks = masks[:, :, top:top + oh, left:left + ow]
    return F.interpolate(masks, tuple(original_size), mode="bilinear", align_corners=False)


def sigmoid_ce_loss(inputs, targets, num_masks):
    loss = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    return loss.flatten(1, 2).mean(1).sum() / (num_masks + 1e-8)


def dice_loss(inputs, targets, eps=1e-6):
    inputs = torch.sigmoid(inputs)
    inputs = inputs.view(inputs.size(0), -1)
    targets = targets.view(targets.size(0), -1)
    inter = (inputs * targets).sum(-1)
    union = inputs.sum(-1) + targets.sum(-1)
    return (1 - (2.0 * inter + eps) / (union + eps)).mean()


def mask_iou_loss(pred_mask, gt, pred_iou):
    p = torch.sigmoid(pred_mask)
    inter = torch.sum(p * gt)
    union = torch.sum(p) + torch.sum(gt) - inter
    iou = (inter + 1e-7) / (union + 1e-7)
    return torch.mean((iou - pred_iou) ** 2)


def focal_loss(pred, mask, gamma=2.0, alpha=0.25):
    p = torch.sigmoid(pred)
    num_pos = torch.sum(mask)
    num_neg = mask.numel() - num_pos
    loss_pos = -alpha * mask * (1 - p) ** gamma * torch.log(p + 1e-12)
    loss_neg = -(1 - alpha) * (1 - mask) * p ** gamma * torch.log(1 - p + 1e-12)
    return (torch.sum(loss_pos) + torch.sum(loss_neg)) / (num_pos + num_neg + 1e-12)


def mask_losses(pred_masks, gt_masks, pred_ious, ce_loss, w):
    """Loss assembly MedPLIB.py:515-572. w: dict(ce, bce, dice, iou, focal) loss weights; ce_loss already weighted."""
    bce = dice = iou = focal = 0
    n = 0
    for i in range(len(pred_masks)):
        gt = gt_masks[i].unsqueeze(0)
        pm = pred_masks[i]
        k = gt.shape[0]
        bce = bce + sigmoid_ce_loss(pm, gt, num_masks=k) * k
        dice = dice + dice_loss(pm, gt) * k
        iou = iou + mask_iou_loss(pm, gt, pred_ious[i]) * k
        focal = focal + focal_loss(pm, gt) * k
        n += k
    u_bce, u_dice, u_iou, u_focal = (x / (n + 1e-8) for x in (bce, dice, iou, focal))
    m_bce, m_dice, m_iou, m_focal = w["bce"] * u_bce, w["dice"] * u_dice, w["iou"] * u_iou, w["focal"] * u_focal
    mask_loss = m_bce + m_dice + m_iou + m_focal
    return {
        "loss": ce_loss + mask_loss, "ce_loss": ce_loss, "mask_bce_loss": m_bce, "mask_dice_loss": m_dice,
        "mask_loss": mask_loss, "unscale_mask_bce_loss": u_bce, "unscale_mask_dice_loss": u_dice,
        "unscale_mask_loss": u_bce + u_dice + u_iou + u_focal, "unscale_mask_iou_loss": u_iou,
        "unscale_mask_focal_loss": u_focal,
    }
