"""Dev tool: per-layer / per-token error of the native LLaMA-MoE stack against the bf16 oracle for one test case, and a
decode-step probe (CPU enqueue time vs GPU time). Usage: python tests/dev/dev_llama.py [case|probe] ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
bf16 = torch.bfloat16
dev = torch.device("cuda:0")


def case(name="tiny_moe", B=8, T=17):
    import test_stacks_gpu as t
    from medplib_b200 import engine
    from oracle import llama, weights
    cfg = t.LLAMA_CFGS[name]
    sd = weights.llama(cfg, seed=21)
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, T, cfg["hidden_size"], generator=g).to(bf16)
    am = torch.ones(B, T + 4, dtype=torch.bool)
    ref = llama.model_forward(sd, cfg, x, am[:, :T])
    eng = engine.LlamaEngine({k: v.to(dev) for k, v in sd.items()}, cfg)
    cache = eng.new_cache(B, T + 8)
    out = eng.forward(x.to(dev).clone(), cache, kv_mask=am.to(dev), want_hidden_states=True, want_router=True)
    torch.cuda.synchronize()
    for l in range(cfg["num_layers"]):
        got, want = out["hidden_states"][l].cpu().float(), ref["hidden_states"][l].float()
        err = (got - want).abs().amax(-1).reshape(-1)
        lg, lr = out["gate_logits"][l].cpu(), ref["gate_logits"][l]
        flip = (lg.argmax(-1) != lr.argmax(-1))
        margin = (lr[:, 0] - lr[:, 1]).abs() / lr.abs().max()
        print(f"layer {l}: hidden-in err max {err.max():.3e} (scale {want.abs().max():.2e}); router err "
              f"{(lg - lr).abs().max():.3e} scale {lr.abs().max():.2e}; flips {flip.nonzero().flatten().tolist()} "
              f"margins {[round(float(margin[i]), 4) for i in flip.nonzero().flatten()]}")
        worst = err.topk(5)
        print("   worst tokens", worst.indices.tolist(), [f"{v:.3f}" for v in worst.values.tolist()])
    got, want = out["last_hidden_state"].cpu().float(), ref["last_hidden_state"].float()
    err = (got - want).abs().amax(-1).reshape(-1)
    print("last: ", err.topk(8).indices.tolist(), [f"{v:.3f}" for v in err.topk(8).values.tolist()])


def probe(B=1, steps=32):
    import bench
    m = bench.build_model(dev)
    eng = m._llama()
    T = 615
    cache = eng.new_cache(B, T + steps + 8)
    x = torch.randn(B, T, 4096, device=dev).to(bf16)
    eng.forward(x, cache)
    torch.cuda.synchronize()
    xs = torch.randn(B, 1, 4096, device=dev).to(bf16)
    for want_router in (False, True):
        for _ in range(3):
            eng.forward(xs.clone(), cache, want_router=want_router)
        cache.len = T
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            eng.forward(xs.clone(), cache, want_router=want_router)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"B={B} want_router={want_router}: enqueue {1e3 * (t1 - t0) / steps:.3f} ms/step, "
              f"gpu {e0.elapsed_time(e1) / steps:.3f} ms/step, wall {1e3 * (t2 - t0) / steps:.3f} ms/step", flush=True)
        cache.len = T


def timing(B=1, layer=5):
    """Per-phase durations of one layer of the decode kernel (stamps by consumer thread 0 of every CTA)."""
    import ctypes
    import bench
    from medplib_b200 import _lib
    lib = _lib.load()
    m = bench.build_model(dev)
    eng = m._llama()
    T = 615
    cache = eng.new_cache(B, T + 64)
    eng.forward(torch.randn(B, T, 4096, device=dev).to(bf16), cache)
    xs = torch.randn(B, 1, 4096, device=dev).to(bf16)
    for _ in range(3):
        eng.forward(xs.clone(), cache)
    cache.len = T
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(16):
        eng.forward(xs.clone(), cache)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B}: {e0.elapsed_time(e1) / 16:.3f} ms/step over 16 steps (context {T}..{T + 16}); env "
          f"LA={os.environ.get('MPL_DK_LA')} SPEC={os.environ.get('MPL_DK_SPEC')} EVICT={os.environ.get('MPL_DK_EVICT')} LA2={os.environ.get('MPL_DK_LA2')}")
    cache.len = T
    lib.mpl_debug_decode_timing(layer, None)
    eng.forward(xs.clone(), cache)
    buf = (ctypes.c_ulonglong * (160 * 32))()
    lib.mpl_debug_decode_timing(-1, buf)
    full = torch.tensor(list(buf), dtype=torch.float64).view(160, 32)[:148]
    # stamps (consumer thread 0 of every CTA): step s in 0..3 = q,k,v / o-proj / gate,up / down:
    #   4s = step start, 4s+1 = staged (+ routed), 4s+2 = tiles done, 4s+3 (s>0) = before the closing barrier,
    #   19+s (s>0) = after it; step 0: 16/17 around the barrier inside the attention phase, 18 = attention done,
    #   3 = after the barrier that closes the attention; 23/24 = router start / logits+softmax done
    us = lambda a, b: (full[:, b] - full[:, a]) / 1e3
    rows = [("  s0: issue cp.async", 0, 25), ("  s0: ln weights (mbar)", 25, 26), ("  s0: cp.async wait", 26, 27),
            ("  s0: sumsq + sync", 27, 28), ("  s0: normalise", 28, 1), ("  s2: router dots", 23, 29),
            ("  s2: router sync", 29, 30), ("  s2: softmax", 30, 24),
            ("s0 stage (RMSNorm)", 0, 1), ("s0 q,k,v tiles", 1, 2), ("s0 -> barrier", 2, 16), ("bar1", 16, 17),
            ("attention", 17, 18), ("bar2", 18, 3), ("s1 stage", 4, 5), ("s1 o-proj tiles", 5, 6), ("bar3", 7, 20),
            ("s2 stage (RMSNorm)", 8, 23), ("s2 router logits", 23, 24), ("s2 scan + publish", 24, 9),
            ("s2 gate/up tiles", 9, 10), ("bar4", 11, 21), ("s3 down tiles", 13, 14), ("bar5", 15, 22)]
    print(f"B={B} layer {layer}: total {((full[:, 22] - full[:, 0]) / 1e3).mean():.1f} us (per-CTA mean)")
    for n, a, b in rows:
        d = us(a, b)
        print(f"  {n:22s} mean {d.mean():7.2f} us  min {d.min():7.2f}  max {d.max():7.2f}")
    t0 = full[:, 0].min()
    for n, i in (("q,k,v tiles", 2), ("attention", 18), ("o-proj tiles", 6), ("gate/up tiles", 10), ("down tiles", 14)):
        print(f"  {n:14s} done: first CTA at {(full[:, i].min() - t0) / 1e3:7.2f} us, last at {(full[:, i].max() - t0) / 1e3:7.2f} us")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "timing":
        timing(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "probe":
        probe(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    else:
        case(*(sys.argv[2:3] or ["tiny_moe"]), *[int(a) for a in sys.argv[3:5]])
