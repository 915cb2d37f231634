"""Dev: where do the full-width evaluate() hidden states differ from the bf16 oracle? python tests/dev/debug_fullwidth.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
import test_fullwidth_gpu as t
from oracle import pipeline, llama, moe as omoe
bf16 = torch.bfloat16
dev = torch.device("cuda:0")
m, sd, ocfg = t.build_full(dev)
g = torch.Generator().manual_seed(2)
ids = torch.randint(3, 31999, (1, 40), generator=g)
ids[0, 2], ids[0, 3], ids[0, 4] = 32001, -200, 32002
clip_img = torch.randn(1, 3, 336, 336, generator=g).to(bf16)
am = torch.ones_like(ids, dtype=torch.bool)
emb_o, am_o, _ = pipeline.prefill_inputs(sd, ocfg, clip_img, ids, am)
_, _, _, emb, _ = m.prepare_inputs_labels_for_multimodal(ids.to(dev), am.to(dev), None, None, clip_img.to(dev), None, None)
e = (emb.float().cpu() - emb_o.float()).abs().amax(-1)[0]
print("inputs_embeds: max err", e.max().item(), "scale", emb_o.float().abs().max().item(), "worst rows", e.topk(5).indices.tolist())
seen = []
hooks = [mod.register_forward_hook(lambda mod_, i, o: seen.append(o.detach().float().cpu()))
         for n, mod in m.named_modules() if "wg" in n and isinstance(mod, torch.nn.Linear)]
eng = m._llama()
cache = eng.new_cache(1, 700)
x = emb_o.to(dev).clone()          # the ORACLE's embeddings: isolates the decoder stack
out = eng.forward(x, cache, want_hidden_states=True, want_router=True)
torch.cuda.synchronize()
gl = out["gate_logits"].cpu()
with omoe.forced_routing([gl[l].view(-1)[:615 * 2].view(615, 2).argmax(-1) for l in range(2)]):
    ref = llama.model_forward(sd, ocfg["llama"], emb_o, am_o)
for l in range(3):
    got = (out["hidden_states"][l] if l < 2 else out["last_hidden_state"]).float().cpu()[0]
    want = ref["hidden_states"][l if l < 2 else 2].float()[0]
    err = (got - want).abs().amax(-1)
    print(f"hidden[{l}] scale {want.abs().max():.3f} max err {err.max():.4f} median row err {err.median():.4f}; rows > 5x median: "
          f"{(err > 5 * err.median()).sum().item()}; worst rows {err.topk(8).indices.tolist()} {[round(v,3) for v in err.topk(8).values.tolist()]}")
    if l == 2:
        bad = err.topk(3).indices.tolist()
        for r in bad:
            d = (got[r] - want[r]).abs()
            print("   row", r, "row scale", want[r].abs().max().item(), "n elems > 0.1:", (d > 0.1).sum().item(), "argmax col", d.argmax().item(),
                  "got", got[r, d.argmax()].item(), "want", want[r, d.argmax()].item())
