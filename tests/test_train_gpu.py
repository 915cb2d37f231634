"""GPU parity of the train step: losses and EVERY trainable parameter's gradient of
MedPLIBForCausalLM.forward(inference=False) — hand-written backward kernels through the C ABI — against torch.autograd
over the CPU oracle (oracle/train.py) in fp32 on the same bf16-rounded weights; then one optimizer step against
torch.optim.AdamW on the oracle's gradients.

Stated tolerances: losses within 2e-2 relative. A gradient tensor passes when max|got - ref| <= 1e-1 * max|ref|
against the fp32 oracle OR against the bf16 oracle (same weights, eager-bf16 arithmetic + bf16 autograd = what the
reference's --precision bf16 training computes), OR its fp32 error is within 1.5x of the bf16 oracle's own fp32 error:
ReLU units / gates whose pre-activation sits inside the bf16 noise flip between precisions, so single-row MLP gradients
(IoU head, hypernetwork, text_hidden_fcs) differ by 20-35 % between the reference's OWN bf16 and fp32 runs (printed).
Routing decisions (expert index per token) are compared bit-exactly first."""
import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
SEG = 299
W = dict(ce=1.0, bce=2.0, dice=0.5, iou=1.0, focal=1.0)
SFT = "wg,lm_head,embed_tokens,mask_decoder,text_hidden_fcs,region_fea_adapter"
TOL = 1e-1


def build(dev, cf=1.5, aux=0.01, lora_targets="q_proj,v_proj,gate_proj,up_proj,down_proj", sft=SFT, dropout=0.0,
          moe_layers=None, top_k=1, lora_r=8):
    import test_model_gpu as tm
    from medplib_b200 import train
    m, _, ocfg = tm.build(dev, moe_layers=moe_layers)
    m.config.moe["capacity_factor"] = cf
    m.config.moe["top_k_experts"] = top_k
    m.config.moe["router_aux_loss_coef"] = aux
    m.router_aux_loss_coef = aux
    m.ce_loss_weight, m.bce_loss_weight, m.dice_loss_weight = W["ce"], W["bce"], W["dice"]
    m.iou_loss_weight, m.focal_loss_weight = W["iou"], W["focal"]
    names = train.attach_lora(m, r=lora_r, lora_alpha=16, lora_dropout=dropout, target_modules=lora_targets)
    if moe_layers is None and lora_targets.count(",") == 4:
        assert len(names) == 2 * 2 + 2 * 2 * 3
    train.set_trainable(m, sft)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "lora_B" in n:  # peft starts B at zero (dA would be identically 0): test a trained-looking adapter
                p.copy_((torch.randn(p.shape, generator=g) * 0.05).to(p.dtype))
            if "wg.weight" in n:
                # router logits of O(1): gates unsaturated (the router gradient is well conditioned); the batch seed
                # below is one where every token's top-1 margin exceeds the bf16-vs-fp32 activation noise
                p.mul_(0.2)
    m.train()
    sd = {k: v.detach().cpu().float() for k, v in m.state_dict().items()}
    sd.update({k: v.detach().cpu().float() for k, v in m.named_buffers()})
    sd["lora_scaling"] = 16.0 / lora_r  # peft: lora_alpha / r
    ocfg["llama"]["moe"] = dict(m.config.moe)
    return m, sd, ocfg


def batch(B=2, n_text=14, seg=False, pad=False, seed=5, region=False):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 290, (B, n_text), generator=g)
    ids[:, 2] = -200
    labels = ids.clone()
    labels[:, :6] = -100
    am = torch.ones_like(ids, dtype=torch.bool)
    if region:  # one region slot per sample (medplib_arch.py:426-429): -300 after the image, no loss on it
        ids[:, 7] = -300
        labels[:, 7] = -100
    if seg:
        ids[:, 9] = SEG
        labels[:, 9] = SEG
    if pad:
        am[1, -3:] = False
        labels[1, -3:] = -100
    clip_img = torch.randn(B, 3, 56, 56, generator=g).to(bf16)
    sam_img = torch.randn(B, 3, 256, 256, generator=g).to(bf16)
    gts = [(torch.rand(70, 90, generator=g) > 0.6).float() for _ in range(B)]
    if region:
        rm = [[(torch.rand(8, 8, generator=g) > 0.6).float()] for _ in range(B)]
        for r in rm:
            r[0][3, 4] = 1.0
        return ids, labels, am, clip_img, sam_img, gts, rm
    return ids, labels, am, clip_img, sam_img, gts


def oracle_run(sd, ocfg, b, seg_flag, noise, dtype=torch.float32):
    from oracle import train as otrain
    ids, labels, am, clip_img, sam_img, gts = b[:6]
    rm = b[6] if len(b) > 6 else None
    out, aux = otrain.train_losses(sd, ocfg, clip_img.to(dtype), sam_img.to(dtype), ids, labels, am, gts,
                                   [tuple(g.shape) for g in gts], [(256, 256)] * len(gts), SEG, W, seg_flag=seg_flag,
                                   rts_uniforms=noise, region_masks=rm,
                                   valid_region=[True] * len(gts) if rm is not None else None)
    return out, aux


def _relerr(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float()
    scale = max(ref.abs().max().item(), 1e-30)
    return (got - ref).abs().max().item() / scale, scale


def oracle_pair(m, sd, ocfg, b, seg_flag, noise, train_names):
    """The oracle twice on the same (bf16-rounded) weights: fp32 arithmetic (the mathematical reference) and bf16
    arithmetic with bf16 autograd (what the reference's own --precision bf16 run computes, rounding where eager
    PyTorch rounds). Their disagreement is the conditioning of each gradient (ReLU / router decisions that flip under
    bf16 noise) and bounds what any bf16 implementation can be held to."""
    sd16 = {k: ((v.to(bf16) if "wg.weight" not in k else v.detach().clone())
                if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in sd.items()}
    for n in train_names:
        sd[n].requires_grad_(True)
        sd16[n].requires_grad_(True)
    ref, aux_o = oracle_run(sd, ocfg, b, seg_flag, noise)
    ref["loss"].backward()
    ref16, aux16 = oracle_run(sd16, ocfg, b, seg_flag, noise, dtype=bf16)
    ref16["loss"].backward()
    return ref, aux_o, sd16, aux16


def relu_bias_order(n_masks):
    md = "model.visual_model.mask_decoder."
    per_mask = [md + f"transformer.layers.{i}.mlp.lin1.bias" for i in range(2)]
    per_mask += [md + f"output_hypernetworks_mlps.{i}.layers.{j}.bias" for i in range(4) for j in range(2)]
    per_mask += [md + f"iou_prediction_head.layers.{j}.bias" for j in range(2)]
    return ["model.text_hidden_fcs.0.0.bias"] + per_mask * n_masks


def calibrate_relu_margins(m, sd, ocfg, b, noise, rounds=12):
    """ReLU is the one non-smooth op of the grounding head: a unit whose pre-activation lies inside the bf16 noise
    band is ON in one precision and OFF in another, and with one-row MLPs (IoU head, hypernetwork, text_hidden_fcs) a
    single flipped unit moves the max-error of a weight gradient by 20-35 % — between the REFERENCE's own bf16 and fp32
    runs too. To make the gradient comparison decisive the biases in front of every ReLU are nudged (fp32 oracle
    forward, CPU) until no pre-activation is within 5 % of its layer's scale of zero; weights stay random."""
    import torch.nn.functional as F
    from oracle import train as otrain
    order = relu_bias_order(len(b[5]))
    params = dict(m.named_parameters())
    for _ in range(rounds):
        rec, on = [], [False]
        orig_relu, orig_fcs, orig_dec = F.relu, otrain.heads.text_hidden_fcs, otrain.sam.mask_decoder

        def relu(x, *a, **k):
            if on[0]:
                rec.append(x.detach())
            return orig_relu(x, *a, **k)

        def scoped(fn):
            def w(*a, **k):
                on[0] = True
                try:
                    return fn(*a, **k)
                finally:
                    on[0] = False
            return w

        F.relu, otrain.heads.text_hidden_fcs, otrain.sam.mask_decoder = relu, scoped(orig_fcs), scoped(orig_dec)
        try:
            with torch.no_grad():
                oracle_run(sd, ocfg, b, True, noise)
        finally:
            F.relu, otrain.heads.text_hidden_fcs, otrain.sam.mask_decoder = orig_relu, orig_fcs, orig_dec
        assert len(rec) == len(order), (len(rec), len(order))
        changed = False
        for name, x in zip(order, rec):
            x2 = x.reshape(-1, x.shape[-1])
            thr = 0.05 * x2.abs().max()
            near = (x2.abs() < thr).any(0)
            if bool(near.any()):
                newb = sd[name].detach().clone()
                newb[near] += 3.0 * thr
                newb = newb.to(bf16).float()
                sd[name] = newb
                with torch.no_grad():
                    params[name].copy_(newb.to(params[name].dtype))
                changed = True
        if not changed:
            return
    raise AssertionError("ReLU margins did not converge")


def dropout_masks(m, S, C, p, seed=21):
    """One keep mask per adapted Linear, [rows, in_features]: rows = tokens for q / v, expert capacity for the MLPs."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, mod in m.named_modules():
        if hasattr(mod, "lora_A"):
            rows = C if "deepspeed_experts" in name else S
            out[name] = torch.rand(rows, mod.in_features, generator=g) >= p
    return out


def run_case(dev, seg_flag, cf, pad, aux, region=False, dropout=0.0, seed=None, **build_kw):
    m, sd, ocfg = build(dev, cf=cf, aux=aux, dropout=dropout, **build_kw)
    if seed is None:
        seed = 27 if region else (205 if dropout > 0 else (8 if build_kw.get("moe_layers") is not None else 5))
    b = batch(seg=seg_flag, pad=pad, region=region, seed=seed)
    ids, labels, am, clip_img, sam_img, gts = b[:6]
    rm = [[x.to(dev) for x in r] for r in b[6]] if region else None
    S = ids.shape[0] * (ids.shape[1] - 1 + 16)
    g = torch.Generator().manual_seed(11)
    noise = [torch.rand(S, 2, generator=g) for _ in range(2)]
    if build_kw.get("moe_layers") is not None:  # oracle / kernels index the noise by decoder layer
        noise = [noise[i] if i in build_kw["moe_layers"] else None for i in range(2)]
    if seg_flag:
        calibrate_relu_margins(m, sd, ocfg, b, noise)
    train_names = [n for n, p in m.named_parameters() if p.requires_grad]
    masks = None
    if dropout > 0:
        from oracle import moe as omoe
        masks = dropout_masks(m, S, omoe.capacity(S, 2, cf, 0), dropout)
        sd["lora_dropout_p"] = dropout
        for k, v in masks.items():
            sd[k + ".lora_dropout_mask"] = v
    ref, aux_o, sd16, aux16 = oracle_pair(m, sd, ocfg, b, seg_flag, noise, train_names)
    tr = m.trainer(lr=1e-2)
    tr.zero_grad()
    tr.stack.dropout_masks = {k: v.to(dev) for k, v in masks.items()} if masks else None
    out = m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), region_masks=rm,
            valid_region_masks_bool=[[True]] * len(gts) if region else None,
            labels=labels.to(dev), attention_mask=am.to(dev), offset=None, masks_list=[x.to(dev) for x in gts],
            label_list=[x.to(dev) for x in gts], resize_list=[(256, 256)] * len(gts), inference=False,
            seg_flag=seg_flag, moe_noise=[x.to(dev) if x is not None else None for x in noise])
    assert set(out) == set(ref)
    # routing must agree exactly, otherwise gradients are not comparable
    assert len(tr.last_gate_logits) == len(aux_o["gate_logits"])
    for l, lg in enumerate(tr.last_gate_logits):
        assert torch.equal(lg.argmax(-1).cpu(), aux_o["gate_logits"][l].argmax(-1)), f"routing differs in layer {l}"
        assert torch.equal(lg.argmax(-1).cpu(), aux16["gate_logits"][l].argmax(-1)), f"routing differs in layer {l}"
    out["loss"].backward()
    torch.cuda.synchronize()
    for k in ref:
        r, o = float(ref[k].detach()), float(out[k].detach())
        assert abs(o - r) <= 2e-2 * max(abs(r), 1e-3) + 1e-4, f"{k}: {o} vs oracle {r}"
    grads = tr.arena.grads()
    assert set(grads) == set(train_names)
    report, bad = [], []
    for n in train_names:
        r32, r16 = sd[n].grad, sd16[n].grad
        if r32 is None:
            assert grads[n].abs().max().item() == 0, f"{n}: gradient should be exactly zero"
            continue
        if r32.abs().max() < 1e-6:  # mathematically zero (e.g. k_proj.bias under softmax shift invariance)
            assert grads[n].abs().max().item() < 1e-3, f"{n}: gradient should vanish"
            continue
        e32, scale = _relerr(grads[n], r32)
        e16, _ = _relerr(grads[n], r16)
        cond, _ = _relerr(r16, r32)
        ok = e16 <= TOL or e32 <= TOL or e32 <= 1.5 * cond
        report.append((min(e16, e32), n, e32, e16, cond, scale))
        if not ok:
            bad.append(n)
    report.sort(reverse=True)
    print("\nworst gradients: min(err16, err32) | name | vs fp32 oracle | vs bf16 oracle | bf16-vs-fp32 oracle | scale")
    for r in report[:14]:
        print("  %.3e  %s  %.3e  %.3e  %.3e  %.3e" % r)
    assert not bad, f"gradients out of tolerance: {bad[:8]} ({len(bad)} of {len(train_names)})"
    return m, tr, sd, train_names


@pytest.mark.parametrize("cf,pad,aux", [(1.0, False, 0.01), (0.4, True, 0.01)])
def test_top2_gating_loss_and_gradients(dev, cf, pad, aux):
    """top_k_experts = 2 (the reference config class's default, model/MedPLIB.py:258): both choices dispatched, combine
    weights renormalised by g1 + g2, capacity 2 * cf * S / E with position-order drops (cf = 0.4 drops second choices and
    some first ones), the injected per-layer noise is top2gating's Gumbel term. Gradients through the renormalisation,
    the softmax and the aux loss vs torch.autograd over oracle/moe.py::top2gating."""
    # (a batch seed where every token's router margin (> 0.12 on the GPU and in both oracles, tests/dev/debug_top2.py) is
    # well above the bf16 activation noise of ~0.04 in both layers)
    run_case(dev, False, cf, pad, aux, top_k=2, seed=53)


@pytest.mark.parametrize("cf,pad,aux", [(1.5, False, 0.01), (0.6, True, 0.0)])
def test_text_loss_and_gradients(dev, cf, pad, aux):
    """seg_flag=False: CE (+ aux) loss; LoRA q,v,gate,up,down of every expert, wg, lm_head, embed_tokens gradients.
    cf=0.6 forces capacity overflow (tokens dropped by injected RTS uniforms); pad=True adds key padding. (The drop case
    uses the batch seed whose smallest router margin, 0.17 over both layers and all three arithmetics, is far above the
    bf16 activation noise of ~0.04 -- tests/dev/debug_top2.py; the default seed sat at 0.014 in layer 1.)"""
    run_case(dev, False, cf, pad, aux, seed=53 if pad else None)


def test_grounding_loss_and_gradients(dev):
    """seg_flag=True: + BCE / Dice / IoU / Focal on the decoded masks; gradients reach the mask decoder,
    text_hidden_fcs and, through the [SEG] hidden rows, the decoder stack."""
    run_case(dev, True, 1.5, False, 0.01)


def test_region_prompt_gradients(dev):
    """rp_flag batches (medplib_arch.py:426-429,580-614): a -300 slot per sample filled with the point-sampled
    region_fea_adapter feature; the adapter's weight / bias gradients come from the input-embedding gradient at the
    slot (sample-mean and Linear commute)."""
    m, tr, sd, names = run_case(dev, False, 1.5, False, 0.0, region=True)
    g = tr.arena.grads()
    assert float(g["model.region_fea_adapter.weight"].abs().max()) > 0
    assert float(sd["model.region_fea_adapter.weight"].grad.abs().max()) > 0


def test_stage2_recipe_dense_layer_all_lora_targets_and_norm_weights(dev):
    """scripts/train_stage2.sh / train_stage3.sh shapes: a DENSE decoder layer next to a MoE layer ("sparse" moe_mode),
    LoRA on all seven projections (q,k,v,o,gate,up,down) and the RMSNorm weights trainable (--sft_modules
    input_layernorm,post_attention_layernorm,mm_projector): exercises the k_proj / o_proj adapter paths, the plain-MLP
    backward, the norm-weight gradients and the projector's gradients (through the image rows of the splice)."""
    run_case(dev, True, 1.5, False, 0.01, moe_layers=[1],
             lora_targets="q_proj,k_proj,v_proj,o_proj,gate_proj,up_proj,down_proj",
             sft="lm_head,embed_tokens,input_layernorm,post_attention_layernorm,model.norm,wg,mask_decoder,"
                 "text_hidden_fcs,mm_projector")


def test_stage2_recipe_rank_16(dev):
    """scripts/train_stage2.sh trains with --lora_r 16 (lora_alpha 16: scaling 1.0) on all seven projections: the
    rank-16 adapters take the separate-pass kernels (the GEMM-fused forms are rank-8 only), same losses and gradients."""
    m, tr, sd, names = run_case(dev, True, 1.5, False, 0.01, moe_layers=[1], lora_r=16,
                                lora_targets="q_proj,k_proj,v_proj,o_proj,gate_proj,up_proj,down_proj",
                                sft="lm_head,embed_tokens,input_layernorm,post_attention_layernorm,model.norm,wg,"
                                    "mask_decoder,text_hidden_fcs,mm_projector")
    assert any(p.shape[0] == 16 for n, p in m.named_parameters() if "lora_A" in n)


def test_lora_dropout(dev):
    """peft's lora_dropout (scripts/train_stage4.sh uses 0.05) with the keep masks injected on both sides: forward
    losses and every gradient, including the masked adapter-input gradient."""
    m, tr, sd, names = run_case(dev, False, 1.5, False, 0.0, dropout=0.25)
    # and it is really on: without the masks the loss differs
    assert tr.stack.training


def test_optimizer_step_matches_adamw(dev):
    """Trainer.step(): global-norm clipping at 1.0 + AdamW(0.9, 0.95) on fp32 masters vs torch.optim.AdamW fed the
    oracle's gradients."""
    m, tr, sd, names = run_case(dev, False, 1.5, False, 0.0)
    params = [sd[n] for n in names]
    before = {n: sd[n].detach().clone() for n in names}
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt = torch.optim.AdamW(params, lr=1e-2, betas=(0.9, 0.95), weight_decay=0.0, eps=1e-8)
    opt.step()
    tr.step()
    torch.cuda.synchronize()
    got = dict(m.named_parameters())
    worst = 0.0
    for n in names:
        delta_ref = (sd[n].detach() - before[n])
        delta = got[n].detach().float().cpu() - before[n]
        if delta_ref.abs().max() == 0:
            continue
        # Adam's first step moves every weight by ~lr * sign(g): compare the update where the gradient is not noise
        big = sd[n].grad.abs() > 0.2 * sd[n].grad.abs().max()
        err = ((delta - delta_ref).abs()[big]).max().item()
        tol = 0.25 * 1e-2 + (2 ** -8) * before[n].abs().max().item()  # + one bf16 ulp of the stored parameter
        worst = max(worst, err / tol)
        assert err <= tol, f"{n}: update differs by {err:.3e} (tol {tol:.3e})"
    assert float(tr.arena.flat.abs().max()) == 0.0  # arena zeroed for the next step
    print(f"\nworst update error / tolerance: {worst:.3f}")


def test_foreign_optimizer_sees_ordinary_grads(dev):
    """Engine expectations of SURVEY 8b (deepspeed.initialize wraps the module, engine.backward = loss.backward, ZeRO
    hangs hooks on the parameters): with Trainer(foreign_grads=True) a plain loss.backward() leaves every trainable
    parameter with an ordinary .grad in its own dtype, delivered through autograd (post-accumulate hooks fire once per
    parameter per backward, .grad accumulates over two backward calls), equal to what the arena path computes; a
    torch.optim step on them is picked up by the next forward."""
    m, sd, ocfg = build(dev)
    b = batch(seg=True)
    m.train()

    def fwd():
        ids, labels, am, clip_img, sam_img, gts = b
        return m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), labels=labels.to(dev),
                 attention_mask=am.to(dev), offset=None, masks_list=[g.to(dev) for g in gts],
                 label_list=[g.to(dev) for g in gts], resize_list=[(256, 256)] * len(gts), inference=False, seg_flag=True,
                 region_masks=None)

    tr = m.trainer()
    fwd()["loss"].backward()
    want = {n: g.clone() for n, g in tr.arena.grads().items()}
    tr.zero_grad()
    tr = m.trainer(foreign_grads=True)
    params = dict(m.named_parameters())
    fired = {}
    for n in want:
        params[n].register_post_accumulate_grad_hook(lambda p, n=n: fired.__setitem__(n, fired.get(n, 0) + 1))
    loss0 = fwd()["loss"]
    loss0.backward()
    assert set(fired) == set(want) and set(fired.values()) == {1}
    for n, g in want.items():
        p = params[n]
        assert p.grad is not None and p.grad.dtype == p.dtype and p.grad.shape == p.shape
        a, w1 = p.grad.float(), g.to(p.dtype).float()  # (atomic accumulation orders differ from run to run: two passes
        # of the SAME path differ by up to 1.4e-2 of a LoRA gradient's scale, tests/dev/debug_determinism.py -- the fp32
        # dQ / weight-gradient REDs land in a different order and the next bf16 rounding amplifies it; a key bias has a
        # mathematically zero gradient -- softmax is shift invariant -- so its value is rounding noise: absolute floor)
        assert (a - w1).abs().max() <= 2.5e-2 * w1.abs().max() + 1e-4, n
    left = {n: float(g.abs().max()) for n, g in tr.arena.grads().items() if float(g.abs().max()) != 0.0}
    assert not left, f"arena not handed over clean: {left}"  # the next micro-step starts from zero
    fwd()["loss"].backward()  # accumulation is autograd's now
    for n in ("lm_head.weight", "model.text_hidden_fcs.0.2.weight"):
        a, w2 = params[n].grad.float(), 2 * want[n].to(params[n].dtype).float()
        assert (a - w2).abs().max() <= 2e-2 * w2.abs().max() + 1e-4, n
    with pytest.raises(Exception):
        tr.step()
    opt = torch.optim.SGD([params[n] for n in want], lr=1e-3)  # (.grad holds two accumulated backward passes)
    opt.step()
    opt.zero_grad()
    loss1 = fwd()["loss"]
    assert float(loss1) < float(loss0)


def test_gradient_accumulation(dev):
    """Two backward passes on the same batch before step() leave exactly twice the single-pass gradient in the arena
    (every kernel accumulates; the lm_head wgrad GEMM goes through a temporary on later micro-steps)."""
    m, sd, ocfg = build(dev, cf=1.5, aux=0.0)
    ids, labels, am, clip_img, sam_img, gts = batch()
    S = ids.shape[0] * (ids.shape[1] - 1 + 16)
    g = torch.Generator().manual_seed(11)
    noise = [torch.rand(S, 2, generator=g).to(dev) for _ in range(2)]
    tr = m.trainer(lr=1e-2)

    def one():
        out = m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), region_masks=None,
                labels=labels.to(dev), attention_mask=am.to(dev), offset=None, masks_list=[], label_list=[],
                resize_list=[], inference=False, seg_flag=False, moe_noise=noise)
        out["loss"].backward()

    tr.zero_grad()
    one()
    single = tr.arena.flat.clone()
    assert tr.micro_steps == 1
    tr.zero_grad()
    with tr.no_sync():
        one()
    one()
    assert tr.micro_steps == 2
    err = (tr.arena.flat - 2 * single).abs().max().item() / single.abs().max().item()
    assert err < 2e-3, err  # fp32 atomics reorder sums run to run
    tr.step()
    assert tr.micro_steps == 0 and float(tr.arena.flat.abs().max()) == 0.0


def test_inference_with_unmerged_adapters_equals_merged_model(dev):
    """train_ds_medplib.py's validate() calls the peft-wrapped model between optimizer steps: evaluate() / generate()
    with adapters attached must equal the merged model (merge_lora), and must follow the weights after a step."""
    from medplib_b200 import train
    import test_model_gpu as tm
    m, sd, ocfg = build(dev, cf=1.5, aux=0.0)
    m.eval()
    ids, clip_img, sam_img = tm.inputs()
    label = torch.zeros(70, 90)
    forced = {3: SEG}
    out_ids, masks = m.evaluate(clip_img.to(dev), sam_img.to(dev), ids.to(dev), [(256, 256)], [label],
                                max_new_tokens=6, forced_tokens=forced)
    m2, _, _ = build(dev, cf=1.5, aux=0.0)  # deterministic: the same weights and adapters
    m2.eval()
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    m2 = train.merge_lora(m2)
    ref_ids, ref_masks = m2.evaluate(clip_img.to(dev), sam_img.to(dev), ids.to(dev), [(256, 256)], [label],
                                     max_new_tokens=6, forced_tokens=forced)
    assert torch.equal(out_ids, ref_ids)
    err = (masks[0].float() - ref_masks[0].float()).abs().max().item()
    assert err <= 2e-2 * ref_masks[0].float().abs().max().item(), err
    # the adapters did matter (the unadapted base model gives different hidden states)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "lora_B" in n:
                p.zero_()
    m.refresh_trained()
    _, base_masks = m.evaluate(clip_img.to(dev), sam_img.to(dev), ids.to(dev), [(256, 256)], [label],
                               max_new_tokens=6, forced_tokens=forced)
    assert (base_masks[0].float() - masks[0].float()).abs().max().item() > 0


def test_lisa_dense_twin_inference_and_train_step(dev):
    """model/LISA.py::LISAForCausalLM (dense LlamaMLP everywhere, `attention_masks` spelling): single-pass grounding
    forward against the oracle pipeline, then one training forward/backward (LoRA q,v + sft modules) against
    torch.autograd over the oracle."""
    import test_model_gpu as tm
    from medplib_b200 import train
    from medplib_b200.model import LISAForCausalLM
    from medplib_b200.model.config import LlavaConfig
    from oracle import pipeline
    torch.manual_seed(0)
    cfg = LlavaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=2, vocab_size=300, rms_norm_eps=1e-5, max_position_embeddings=512,
                      mm_vision_select_layer=-2, mm_projector_type="mlp2x_gelu", initializer_range=0.06)
    cfg.clip_config = tm.CLIP_CFG
    cfg.sam_config = dict(image_size=256, embed_dim=128, depth=3, num_heads=2)
    m = LISAForCausalLM(cfg, seg_token_idx=SEG, use_mm_start_end=True, train_mask_decoder=True, out_dim=256,
                        ce_loss_weight=W["ce"], dice_loss_weight=W["dice"], bce_loss_weight=W["bce"],
                        iou_loss_weight=W["iou"], focal_loss_weight=W["focal"], vision_tower=None,
                        region_fea_adapter=True, region_geo_sampler=False, max_sample_point=512,
                        sampler_pooler_mode="max")
    assert not any("deepspeed_moe" in n for n, _ in m.named_parameters())
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "rel_pos" in n or "pos_embed" in n or n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.06)
            if "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
    m.config.mm_use_im_start_end = True
    m = m.to(bf16).to(dev).eval()
    ocfg = dict(clip=dict(hidden_size=128, intermediate_size=256, num_layers=3, num_heads=2, image_size=56,
                          patch_size=14),
                llama=dict(hidden_size=256, intermediate_size=512, num_layers=2, num_heads=2, vocab_size=300,
                           rms_norm_eps=1e-5, max_position_embeddings=512, rope_theta=1e4, moe=None),
                sam=dict(num_heads=2), mm_use_im_start_end=True, mm_token_compress=False)
    # ---- inference (model_forward(inference=True), LISA's `attention_masks` keyword)
    sd_b = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sd_b.update({k: v.detach().cpu() for k, v in m.named_buffers()})
    ids, clip_img, sam_img = tm.inputs(seg_in_prompt=True)
    label = torch.zeros(70, 90)
    ref = pipeline.grounding_forward(sd_b, ocfg, clip_img, sam_img, ids, [(256, 256)], [(70, 90)], SEG)
    out = m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), region_masks=None,
            labels=None, attention_masks=torch.ones_like(ids, dtype=torch.bool).to(dev), offset=None,
            masks_list=[label], label_list=[label], resize_list=[(256, 256)], inference=True)
    tm._check(out["pred_masks"][0], ref["pred_masks"][0], 3.5e-2, "LISA mask logits")
    # ---- one train step
    train.attach_lora(m, r=8, lora_alpha=16, target_modules="q_proj,v_proj")
    train.set_trainable(m, "lm_head,embed_tokens,mask_decoder,text_hidden_fcs")
    gg = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "lora_B" in n:
                p.copy_((torch.randn(p.shape, generator=gg) * 0.05).to(p.dtype))
    m.train()
    sd = {k: v.detach().cpu().float() for k, v in m.state_dict().items()}
    sd.update({k: v.detach().cpu().float() for k, v in m.named_buffers()})
    sd["lora_scaling"] = 2.0
    b = batch(seg=True)
    calibrate_relu_margins(m, sd, ocfg, b, None, rounds=40)
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    ref_l, aux_o, sd16, aux16 = oracle_pair(m, sd, ocfg, b, True, None, names)
    ids, labels, am, clip_img, sam_img, gts = b
    tr = m.trainer(lr=1e-2)
    tr.zero_grad()
    out = m(images=sam_img.to(dev), images_clip=clip_img.to(dev), input_ids=ids.to(dev), region_masks=None,
            labels=labels.to(dev), attention_masks=am.to(dev), offset=None, masks_list=[x.to(dev) for x in gts],
            label_list=[x.to(dev) for x in gts], resize_list=[(256, 256)] * len(gts), inference=False, seg_flag=True)
    out["loss"].backward()
    torch.cuda.synchronize()
    for k in ref_l:
        r, o = float(ref_l[k].detach()), float(out[k].detach())
        assert abs(o - r) <= 2e-2 * max(abs(r), 1e-3) + 1e-4, f"{k}: {o} vs oracle {r}"
    grads = tr.arena.grads()
    bad = []
    for n in names:
        r32, r16 = sd[n].grad, sd16[n].grad
        if r32 is None or r32.abs().max() < 1e-6:
            continue
        e32, _ = _relerr(grads[n], r32)
        e16, _ = _relerr(grads[n], r16)
        cond, _ = _relerr(r16, r32)
        if not (e16 <= TOL or e32 <= TOL or e32 <= 1.5 * cond):
            bad.append((n, e32, e16))
    assert not bad, bad[:6]


def test_icl_recipe_compressor_and_mask_encoder_gradients(dev):
    """scripts/train_medplib_icl.sh with ICL_MASK_MODE=separate: --sft_modules mask_decoder,text_hidden_fcs,mask_encoder,
    mm_token_compressor, LoRA on gate/up/down, two (image, mask) exemplars + the query image per sample, compressed
    image tokens (16 -> 8) and MaskTokenEncoder tokens (4 per mask), [SEG] + ground-truth mask. Every trainable
    tensor's gradient -- the TokenCompressor's LayerNorm / Linear, the MaskTokenEncoder's four convs, Linear and
    LayerNorm included -- against torch.autograd over oracle/train.py::train_losses_icl; mm_projector trained too in a
    second pass (its gradient flows back through the pool)."""
    import test_model_gpu as tm
    from medplib_b200 import train
    from oracle import train as otrain
    for sft in ("mask_decoder,text_hidden_fcs,mask_encoder,mm_token_compressor",
                "mask_encoder,mm_token_compressor,mm_projector"):
        m, _, ocfg = tm.build_icl(dev)
        m.config.moe["capacity_factor"] = 1.5
        m.router_aux_loss_coef = 0.01
        m.ce_loss_weight, m.bce_loss_weight, m.dice_loss_weight = W["ce"], W["bce"], W["dice"]
        m.iou_loss_weight, m.focal_loss_weight = W["iou"], W["focal"]
        train.attach_lora(m, r=8, lora_alpha=16, lora_dropout=0.0, target_modules="gate_proj,up_proj,down_proj")
        train.set_trainable(m, sft)
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "lora_B" in n:
                    p.copy_((torch.randn(p.shape, generator=g) * 0.05).to(p.dtype))
                if "wg.weight" in n:
                    p.mul_(0.2)
        m.train()
        sd = {k: v.detach().cpu().float() for k, v in m.state_dict().items()}
        sd.update({k: v.detach().cpu().float() for k, v in m.named_buffers()})
        sd["lora_scaling"] = 2.0
        ocfg["llama"]["moe"] = dict(m.config.moe)
        g = torch.Generator().manual_seed(4)
        types_ = [["image", "mask", "image", "mask", "image"]]
        lengths = [[8, 4, 8, 4, 8]]
        ids = torch.randint(3, 290, (1, 26), generator=g)
        for pos in (2, 6, 10, 14, 18):
            ids[0, pos] = -200
        ids[0, 23] = SEG
        labels = ids.clone()
        labels[:, :20] = -100
        am = torch.ones_like(ids, dtype=torch.bool)
        clip_imgs = [torch.randn(3, 3, 56, 56, generator=g).to(bf16)]
        mask_imgs = [(torch.rand(2, 1, 56, 56, generator=g) > 0.8).to(bf16)]
        sam_img = torch.randn(1, 3, 256, 256, generator=g).to(bf16)
        gts = [(torch.rand(70, 90, generator=g) > 0.6).float()]
        S = 26 - 5 + 3 * 8 + 2 * 4
        noise = [torch.rand(S, 2, generator=g) for _ in range(2)]
        train_names = [n for n, p in m.named_parameters() if p.requires_grad]
        assert any("mm_token_compressor" in n for n in train_names) and any("mask_encoder.encoder.0" in n for n in train_names)

        def oracle(sdx, dtype):
            out, aux = otrain.train_losses_icl(sdx, ocfg, [c.to(dtype) for c in clip_imgs], [x.to(dtype) for x in mask_imgs],
                                               types_, lengths, sam_img.to(dtype), ids, labels, am, gts, [(70, 90)],
                                               [(256, 256)], SEG, W, rts_uniforms=noise)
            return out, aux

        sd16 = {k: ((v.to(bf16) if "wg.weight" not in k else v.detach().clone())
                    if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in sd.items()}
        for n in train_names:
            sd[n].requires_grad_(True)
            sd16[n].requires_grad_(True)
        ref, aux_o = oracle(sd, torch.float32)
        ref["loss"].backward()
        ref16, _ = oracle(sd16, bf16)
        ref16["loss"].backward()
        tr = m.trainer(lr=1e-2)
        tr.zero_grad()
        out = m(images=sam_img.to(dev), images_clip=[c.to(dev) for c in clip_imgs], input_ids=ids.to(dev),
                region_masks=None, labels=labels.to(dev), attention_mask=am.to(dev), offset=None,
                masks_list=[x.to(dev) for x in gts], label_list=[x.to(dev) for x in gts], resize_list=[(256, 256)],
                inference=False, mask_images=[x.to(dev) for x in mask_imgs], image_token_types=types_,
                image_token_lengths=lengths, icl_image_counts=[3], moe_noise=[x.to(dev) for x in noise])
        for l, lg in enumerate(tr.last_gate_logits):
            assert torch.equal(lg.argmax(-1).cpu(), aux_o["gate_logits"][l].argmax(-1)), f"routing differs in layer {l}"
        out["loss"].backward()
        torch.cuda.synchronize()
        for k in ref:
            r, o = float(ref[k].detach()), float(out[k].detach())
            assert abs(o - r) <= 3e-2 * max(abs(r), 1e-3) + 1e-4, f"{k}: {o} vs oracle {r}"
        grads = tr.arena.grads()
        assert set(grads) == set(train_names)
        bad, report = [], []
        for n in train_names:
            if not ("mm_token_compressor" in n or "mask_encoder" in n or "mm_projector" in n):
                continue  # (the rest of the step is covered by the tests above)
            r32, r16 = sd[n].grad, sd16[n].grad
            e32, scale = _relerr(grads[n], r32)
            e16, _ = _relerr(grads[n], r16)
            cond, _ = _relerr(r16, r32)
            report.append((n, e32, e16, cond, scale))
            if not (e16 <= TOL or e32 <= TOL or e32 <= 1.5 * cond):
                bad.append(n)
        print("\nvision-side adapter gradients: name | vs fp32 oracle | vs bf16 oracle | bf16-vs-fp32 oracle | scale")
        for r in report:
            print("  %s  %.3e  %.3e  %.3e  %.3e" % r)
        assert len(report) >= 12 and not bad, f"gradients out of tolerance: {bad}"
