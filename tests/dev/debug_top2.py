"""Dev probe (GPU box): router margins of a train-test case, GPU vs fp32 / bf16 oracle, over a few batch seeds.
   python tests/dev/debug_top2.py [cf pad(0/1) top_k seeds,comma,separated]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import test_train_gpu as tt
dev = torch.device("cuda:0")
cf, pad, top_k = (float(sys.argv[1]), sys.argv[2] == "1", int(sys.argv[3])) if len(sys.argv) > 3 else (0.4, True, 2)
seeds = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else (52, 53, 5, 8, 21, 33, 47)
for seed in seeds:
    m, sd, ocfg = tt.build(dev, cf=cf, aux=0.0 if top_k == 1 else 0.01, top_k=top_k)
    b = tt.batch(seg=False, pad=pad, seed=seed)
    ids, labels, am, clip_img, sam_img, gts = b[:6]
    S = ids.shape[0] * (ids.shape[1] - 1 + 16)
    g = torch.Generator().manual_seed(11)
    noise = [torch.rand(S, 2, generator=g) for _ in range(2)]
    with torch.no_grad():
        _, aux_o = tt.oracle_run(sd, ocfg, b, False, noise)
        sd16 = {k: ((v.to(torch.bfloat16) if "wg.weight" not in k else v.detach().clone())
                    if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in sd.items()}
        _, aux16 = tt.oracle_run(sd16, ocfg, b, False, noise, dtype=torch.bfloat16)
    tr = m.trainer(lr=1e-2)
    tr.zero_grad()
    out = m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), region_masks=None,
            valid_region_masks_bool=None, labels=labels.to(dev), attention_mask=am.to(dev), offset=None,
            masks_list=[x.to(dev) for x in gts], label_list=[x.to(dev) for x in gts], resize_list=[(256, 256)] * len(gts),
            inference=False, seg_flag=False, moe_noise=[x.to(dev) for x in noise])
    for l, lg in enumerate(tr.last_gate_logits):
        lg = lg.cpu().float()
        o32, o16 = aux_o["gate_logits"][l].float(), aux16["gate_logits"][l].float()
        bad = (lg.argmax(-1) != o32.argmax(-1)) | (lg.argmax(-1) != o16.argmax(-1))
        print(f"seed {seed} layer {l}: mismatches {int(bad.sum())}; min margin gpu {float((lg[:,0]-lg[:,1]).abs().min()):.3f} "
              f"fp32 {float((o32[:,0]-o32[:,1]).abs().min()):.3f} bf16 {float((o16[:,0]-o16[:,1]).abs().min()):.3f}; "
              f"max |gpu - fp32| {float((lg-o32).abs().max()):.3f} max |bf16 - fp32| {float((o16-o32).abs().max()):.3f}")
        for i in torch.nonzero(bad).flatten().tolist():
            print("   token", i, "gpu", lg[i].tolist(), "fp32", o32[i].tolist(), "bf16", o16[i].tolist())
