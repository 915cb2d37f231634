"""GPU-side localisation of grounding-head gradient mismatches (forward states + gradients at the head's boundaries)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_train_gpu as tt  # noqa: E402
from medplib_b200 import mask_train  # noqa: E402

dev = torch.device("cuda:0")
m, sd, ocfg = tt.build(dev, cf=1.5, aux=0.0)
b = tt.batch(seg=True, pad=False)
ids, labels, am, clip_img, sam_img, gts = b
S = ids.shape[0] * (ids.shape[1] - 1 + 16)
g = torch.Generator().manual_seed(11)
noise = [torch.rand(S, 2, generator=g) for _ in range(2)]
names = [n for n, p in m.named_parameters() if p.requires_grad]
for n in names:
    sd[n].requires_grad_(True)
ref, aux = tt.oracle_run(sd, ocfg, b, True, noise)
aux["pred_embeddings"].retain_grad()
for t in aux["low_res"] + aux["pred_masks"] + aux["pred_ious"]:
    t.retain_grad()
ref["loss"].backward()

cap = {}
orig_head = mask_train.mask_head_losses
orig_dec = mask_train.mask_decoder


def dec(tr, model, img_tok, text):
    low, iou = orig_dec(tr, model, img_tok, text)
    i = len(cap.setdefault("low", []))
    cap["low"].append(low.detach().clone())
    cap.setdefault("iou", []).append(iou.detach().clone())
    low.register_hook(lambda gr, i=i: cap.__setitem__(f"dlow{i}", gr.detach().clone()))
    iou.register_hook(lambda gr, i=i: cap.__setitem__(f"diou{i}", gr.detach().clone()))
    return low, iou


def head(tr, model, pe, *a):
    cap["pe"] = pe.detach().clone()
    pe.register_hook(lambda gr: cap.__setitem__("dpe", gr.detach().clone()))
    out = orig_head(tr, model, pe, *a)
    for i, pm in enumerate(out["pred_masks"]):
        cap[f"pm{i}"] = pm.detach().clone()
        pm.register_hook(lambda gr, i=i: cap.__setitem__(f"dpm{i}", gr.detach().clone()))
    return out


mask_train.mask_decoder = dec
mask_train.mask_head_losses = head
tr = m.trainer(lr=1e-2)
tr.zero_grad()
out = m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), region_masks=None,
        labels=labels.to(dev), attention_mask=am.to(dev), offset=None, masks_list=[x.to(dev) for x in gts],
        label_list=[x.to(dev) for x in gts], resize_list=[(256, 256)] * len(gts), inference=False, seg_flag=True,
        moe_noise=[x.to(dev) for x in noise])
out["loss"].backward()
torch.cuda.synchronize()


def rel(a, r):
    a, r = a.detach().float().cpu().reshape(r.shape), r.detach().float()
    return "%.3e (scale %.3e)" % ((a - r).abs().max().item() / max(r.abs().max().item(), 1e-12), r.abs().max().item())


for k in ref:
    print(k, float(out[k]), float(ref[k]))
print("pred_embeddings fwd", rel(cap["pe"], aux["pred_embeddings"]))
print("d pred_embeddings  ", rel(cap["dpe"], aux["pred_embeddings"].grad))
for i in range(len(gts)):
    print(f"mask {i}: low fwd", rel(cap["low"][i], aux["low_res"][i]), "| pm fwd", rel(cap[f"pm{i}"], aux["pred_masks"][i]),
          "| iou fwd", float(cap["iou"][i]), float(aux["pred_ious"][i]))
    print(f"   d pm ", rel(cap[f"dpm{i}"], aux["pred_masks"][i].grad), "| d low", rel(cap[f"dlow{i}"], aux["low_res"][i].grad),
          "| d iou", float(cap[f"diou{i}"]), float(aux["pred_ious"][i].grad))
    p = torch.sigmoid(aux["pred_masks"][i].detach())
    I = (p * gts[i]).sum(); U = p.sum() + gts[i].sum() - I
    print("   oracle iou", float(I / U))
grads = tr.arena.grads()
rows = []
for n in names:
    if "visual_model" in n or "text_hidden" in n:
        rg = sd[n].grad
        if rg is None or rg.abs().max() < 1e-7:
            continue
        e = (grads[n].cpu() - rg).abs().max().item() / rg.abs().max().item()
        rows.append((e, n, rg.abs().max().item()))
rows.sort(reverse=True)
for e, n, s in rows[:40]:
    print("%.3e %s %.3e" % (e, n, s))
