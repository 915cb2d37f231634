"""CPU tests of the train step's host logic (no kernels run here): LoRA attachment mirrors peft's module surgery and the
reference's target matching, the gradient arena's layout / bucket order, the data-parallel bucket reducer over gloo at
world_size 2, and self-consistency of the training oracle (LoRA == merged weights; autograd reaches every trainable)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
bf16 = torch.bfloat16


def _model():
    import test_train_gpu as tt
    return tt.build(torch.device("cpu"))


def test_attach_lora_and_sft_modules_follow_the_reference_matching():
    from medplib_b200 import train
    m, sd, _ = _model()
    names = [n for n, _ in m.named_parameters()]
    lora = [n for n in names if "lora_" in n]
    # peft's parameter names under the matched Linear; q,v of 2 layers + gate,up,down of 2 experts x 2 layers
    assert "model.layers.0.self_attn.q_proj.lora_A.default.weight" in lora
    assert "model.layers.1.mlp.deepspeed_moe.experts.deepspeed_experts.1.down_proj.lora_B.default.weight" in lora
    assert len(lora) == 2 * (2 * 2 + 2 * 2 * 3)
    assert not any("k_proj.lora" in n or "o_proj.lora" in n for n in lora)
    assert not any(x in n for n in lora for x in train.LORA_EXCLUDE)  # train_ds_medplib.py:272-281
    p = dict(m.named_parameters())
    assert p["model.layers.0.self_attn.q_proj.weight"].requires_grad is False
    assert p["model.layers.0.self_attn.q_proj.lora_A.default.weight"].shape == (8, 256)
    for n in ("lm_head.weight", "model.embed_tokens.weight", "model.layers.0.mlp.deepspeed_moe.gate.wg.weight",
              "model.text_hidden_fcs.0.0.weight", "model.visual_model.mask_decoder.iou_token.weight",
              "model.region_fea_adapter.weight"):
        assert p[n].requires_grad, n
    assert not p["model.visual_model.image_encoder.pos_embed"].requires_grad
    assert not p["model.mm_projector.0.weight"].requires_grad
    assert p["model.layers.0.mlp.deepspeed_moe.gate.wg.weight"].dtype == torch.float32
    # no CPU fallback for the train step either
    from medplib_b200 import _lib
    with pytest.raises(_lib.MplError):
        m.trainer()


def test_arena_layout_is_backward_completion_order():
    from medplib_b200 import train
    m, _, _ = _model()
    named = [(n, p) for n, p in m.named_parameters() if p.requires_grad]
    ordered = train.Trainer._order(m, named)
    arena = train.GradArena(ordered, torch.device("cpu"))
    pos = {n: o for n, o in zip(arena.names, arena.offsets)}
    assert pos["lm_head.weight"] < pos["model.layers.1.self_attn.q_proj.lora_A.default.weight"] \
        < pos["model.layers.0.self_attn.q_proj.lora_A.default.weight"] < pos["model.embed_tokens.weight"]
    assert pos["model.visual_model.mask_decoder.iou_token.weight"] < pos["lm_head.weight"]
    assert all(o % train.GradArena.ALIGN == 0 for o in arena.offsets)
    p = dict(named)["lm_head.weight"]
    v = arena.of(p)
    assert v.shape == p.shape and v.dtype == torch.float32
    v.fill_(2.0)
    assert float(arena.flat.sum()) == 2.0 * p.numel()
    assert arena.of(dict(m.named_parameters())["model.layers.0.self_attn.q_proj.weight"]) is None
    arena.export_grads()
    assert p.grad.dtype == p.dtype and float(p.grad.float().mean()) == 2.0


def _reducer_worker(rank, world, port, q):
    import torch.distributed as dist
    from medplib_b200 import train
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ps = [("a", torch.nn.Parameter(torch.zeros(100))), ("b", torch.nn.Parameter(torch.zeros(7, 9))),
          ("c", torch.nn.Parameter(torch.zeros(300)))]
    arena = train.GradArena(ps, torch.device("cpu"))
    red = train.BucketReducer(arena, bucket_elems=128)
    assert red.on and red.world == world and len(red.bounds) == (arena.numel + 127) // 128
    for _, p in ps:
        arena.of(p).fill_(float(rank + 1))
    red.ready(arena.end_of(ps[0][1]))  # first parameter done: no full bucket yet (100 < 128)
    n_early = red.next
    red.ready(arena.end_of(ps[1][1]))
    n_mid = red.next
    scale = red.finish()
    total = sum(range(1, world + 1))
    ok = all(torch.allclose(arena.of(p) * scale, torch.full(p.shape, total / world)) for _, p in ps)
    q.put((rank, n_early, n_mid, ok, scale))
    dist.destroy_process_group()


def test_bucket_reducer_gloo_world2():
    """N>1 path of the train step: every rank ends with the mean gradient; buckets launch as soon as the
    backward has passed their end."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for rank, n_early, n_mid, ok, scale in res:
        assert n_early == 0 and n_mid == 1 and ok and scale == 0.5, res


def test_training_oracle_lora_equals_merged_weights_and_reaches_every_trainable():
    """Pins the LoRA restatement by its defining identity: the adapter path == the same model with W + s*B*A merged
    (peft merge_and_unload), and checks autograd reaches every trainable tensor with finite values."""
    import test_train_gpu as tt
    m, sd, ocfg = _model()
    b = tt.batch(seg=True)
    S = b[0].shape[0] * (b[0].shape[1] - 1 + 16)
    g = torch.Generator().manual_seed(11)
    noise = [torch.rand(S, 2, generator=g) for _ in range(2)]
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    for n in names:
        sd[n].requires_grad_(True)
    out, aux = tt.oracle_run(sd, ocfg, b, True, noise)
    assert set(out) == {"loss", "ce_loss", "mask_bce_loss", "mask_dice_loss", "mask_loss", "unscale_mask_bce_loss",
                        "unscale_mask_dice_loss", "unscale_mask_loss", "unscale_mask_iou_loss",
                        "unscale_mask_focal_loss"}
    out["loss"].backward()
    missing = [n for n in names if sd[n].grad is None and "region_fea_adapter" not in n]
    assert not missing, missing
    assert all(torch.isfinite(sd[n].grad).all() for n in names if sd[n].grad is not None)
    merged = {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in sd.items()}
    for k in list(merged):
        if k.endswith(".lora_A.default.weight"):
            base = k[: -len(".lora_A.default.weight")]
            merged[base + ".weight"] = merged[base + ".weight"] + 2.0 * merged[base + ".lora_B.default.weight"] @ merged[k]
            del merged[k], merged[base + ".lora_B.default.weight"]
    with torch.no_grad():
        out2, _ = tt.oracle_run(merged, ocfg, b, True, noise)
    assert abs(float(out2["loss"]) - float(out["loss"])) < 2e-4 * abs(float(out["loss"]))
