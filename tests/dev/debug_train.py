"""GPU-side localisation of a train-step gradient mismatch: per-layer forward states and activation gradients of the
CUDA path (LlamaTrainStack.debug) against the CPU oracle's autograd (retain_grad on the layer inputs)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_train_gpu as tt  # noqa: E402

cf, pad, aux = float(sys.argv[1]), sys.argv[2] == "1", float(sys.argv[3])
dev = torch.device("cuda:0")
m, sd, ocfg = tt.build(dev, cf=cf, aux=aux)
b = tt.batch(seg=False, pad=pad)
ids, labels, am, clip_img, sam_img, gts = b
S = ids.shape[0] * (ids.shape[1] - 1 + 16)
g = torch.Generator().manual_seed(11)
noise = [torch.rand(S, 2, generator=g) for _ in range(2)]
names = [n for n, p in m.named_parameters() if p.requires_grad]
for n in names:
    sd[n].requires_grad_(True)

# oracle with retained grads on every layer input
from oracle import llama  # noqa: E402
orig = llama.decoder_layer
keep = {}


def hooked(sd_, i, x, *a, **k):
    x.retain_grad()
    keep[f"x{i}"] = x
    out = orig(sd_, i, x, *a, **k)
    return out


llama.decoder_layer = hooked
ref, aux_o = tt.oracle_run(sd, ocfg, b, False, noise)
aux_o["hidden"].retain_grad()
ref["loss"].backward()

tr = m.trainer(lr=1e-2)
tr.zero_grad()
tr.stack.debug = {}
out = m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), region_masks=None,
        labels=labels.to(dev), attention_mask=am.to(dev), offset=None, masks_list=[], label_list=[], resize_list=[],
        inference=False, seg_flag=False, moe_noise=[x.to(dev) for x in noise])
out["loss"].backward()
torch.cuda.synchronize()
d = tr.stack.debug


def rel(a, r):
    a, r = a.detach().float().cpu().reshape(r.shape), r.detach().float()
    return (a - r).abs().max().item() / max(r.abs().max().item(), 1e-9)


print("loss", float(out["loss"]), float(ref["loss"]))
print("fwd x1 (layer0 out):", rel(d["f0.out"], keep["x1"]))
print("bwd: d loss/d x2-in(final):", "n/a")
print("bwd dx into layer1 out vs oracle dhidden-chain: b1.dout ~ grad of layer1 output (pre-norm)")
print("b1.dx vs oracle grad x1:", rel(d["b1.dx"], keep["x1"].grad))
print("b0.dx vs oracle grad x0:", rel(d["b0.dx"], keep["x0"].grad))
for k in ("f0.slot", "f1.slot"):
    sl = d[k].cpu()[:, 0]
    print(k, "dropped:", int((sl < 0).sum()), "of", sl.numel())
for l in (1, 0):
    for k in ("dout", "dy", "dxin", "dn2", "dh1", "do", "dqkv", "dn1", "dx"):
        t = d[f"b{l}.{k}"]
        print(f"  b{l}.{k}: absmax {t.float().abs().max().item():.4e} mean|.| {t.float().abs().mean().item():.4e}")
