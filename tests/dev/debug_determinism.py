"""Dev probe (GPU box): two forward+backward passes on the same batch / weights -> which gradients differ, by how much."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import test_train_gpu as tt
dev = torch.device("cuda:0")
m, sd, ocfg = tt.build(dev)
b = tt.batch(seg=True)
m.train()
def fwd():
    ids, labels, am, clip_img, sam_img, gts = b
    return m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), labels=labels.to(dev),
             attention_mask=am.to(dev), offset=None, masks_list=[g.to(dev) for g in gts],
             label_list=[g.to(dev) for g in gts], resize_list=[(256, 256)] * len(gts), inference=False, seg_flag=True,
             region_masks=None)
tr = m.trainer()
runs = []
for _ in range(3):
    tr.zero_grad()
    fwd()["loss"].backward()
    torch.cuda.synchronize()
    runs.append({n: g.clone() for n, g in tr.arena.grads().items()})
worst = []
for n in runs[0]:
    d = max(float((runs[i][n] - runs[0][n]).abs().max()) for i in (1, 2))
    worst.append((d / max(float(runs[0][n].abs().max()), 1e-30), n))
worst.sort(reverse=True)
print("env", {k: v for k, v in os.environ.items() if k.startswith("MPL_")})
for r, n in worst[:8]:
    print(f"  {r:.3e}  {n}")
