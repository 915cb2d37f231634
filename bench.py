"""bench.py — MedPLIB-7B pixel grounding (BASELINE.json configs[1]) on N B200s.

  python bench.py --gpus N --steps K --warmup W              ours: hand-written sm_100a kernels behind the C ABI.
                                                             ONE JSON line: pixel grounding (configs[1]) is the headline;
                                                             "secondary" carries the other BASELINE configs measured in the
                                                             same run: decode (configs[2], tokens/s), train (configs[3],
                                                             samples/s, data parallel over the N ranks), icl (configs[4])
  python bench.py --impl reference --gpus N --steps K ...    the reference's path on the host cores (CPU oracle port)
  python bench.py --workload grounding|decode|train|icl ...  one workload alone (its own line)
  python bench.py --workload preprocess ...                  GPU image input pipeline, images/s

A "step" is one image through MedPLIBForCausalLM.evaluate(): CLIP-L/14-336 -> mm_projector -> splice (T = 40 + 575) ->
LLaMA-7B-MoE (2 experts, top-1) prefill -> 8 greedy decode tokens (<SEG> forced at new token 4, since random weights
never emit it) -> text_hidden_fcs -> SAM-Med2D ViT-B encoder -> two-way mask decoder -> bilinear resize to 336x336.
Weights: random init of the reference architecture (no checkpoints offline), bf16. Inference shards as independent
replicas: with N ranks every rank runs its own images, value = N * K images / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
bf16 = torch.bfloat16
SEG = 32003
N_TEXT, N_NEW, SEG_AT = 40, 8, 4
DIMS = dict(D=4096, F=11008, L=32, H=32, V=32267, E=2)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sus=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


def make_inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    images_clip = torch.randn(1, 3, 336, 336, generator=torch.Generator().manual_seed(0))
    images = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(1))
    ids = torch.randint(3, 31999, (1, N_TEXT), generator=torch.Generator().manual_seed(2))
    ids[0, 2], ids[0, 3], ids[0, 4] = 32001, -200, 32002  # <im_start> IMAGE <im_end>
    return images_clip, images, ids


# ------------------------------------------------------------------------------------------------- our arm
def build_model(dev, small=False, icl=False):
    from medplib_b200.model import MedPLIBForCausalLM, MedPLIBMoELlamaConfig
    d = DIMS if not small else dict(D=512, F=1024, L=2, H=4, V=32267, E=2)
    cfg = MedPLIBMoELlamaConfig(hidden_size=d["D"], intermediate_size=d["F"], num_hidden_layers=d["L"],
                                num_attention_heads=d["H"], num_key_value_heads=d["H"], vocab_size=d["V"],
                                rms_norm_eps=1e-5, max_position_embeddings=4096, mm_vision_select_layer=-2,
                                mm_projector_type="mlp2x_gelu", max_sample_point=512)
    cfg.moe = dict(num_experts=[d["E"]], top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0,
                   min_capacity=0, use_residual=False, router_aux_loss_coef=0.01, moe_layers_idx=None,
                   moe_mode="dense", ep_size=1)
    old = torch.get_default_dtype()
    torch.set_default_dtype(bf16)
    try:
        with torch.device(dev):
            kw = dict(mm_token_compress=True, mm_compressed_token_count=256, icl_mask_encoder=True,
                      mask_encoder_token_count=64, use_mm_start_end=True) if icl else {}
            m = MedPLIBForCausalLM(cfg, test_only=True, seg_token_idx=SEG, **kw)
    finally:
        torch.set_default_dtype(old)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "deepspeed_experts.1." in n:  # experts start as copies (like the reference); make them differ
                p.normal_(0.0, 0.02)
    m.config.eos_token_id = -1  # never stop early: fixed work per step
    m.config.mm_use_im_start_end = True
    return m.to(bf16).to(dev).eval()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.rows, self.p, self.gpu = [], None, gpu

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = max([int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 4 + i and r[4 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


def gemm_flops_per_image(T=N_TEXT - 1 + 576, n_clip=1):
    """Algorithmic FLOPs of the tcgen05 GEMM launches of one step (2*M*N*K; top-1 MoE = dense FFN FLOPs): decoder
    linears over T positions + n_clip x (CLIP tower + projector) + one SAM-Med2D encoder."""
    d = DIMS
    llama = T * d["L"] * (4 * d["D"] * d["D"] + 3 * d["D"] * d["F"]) * 2
    clip = 577 * 23 * (4 * 1024 * 1024 + 2 * 1024 * 4096) * 2 + 576 * 592 * 1024 * 2
    proj = 576 * (1024 * 4096 + 4096 * 4096) * 2
    sam = 12 * (256 * (4 * 768 * 768 + 2 * 768 * 3072) * 2) + 8 * (784 - 256) * 3 * 768 * 768 * 2 \
        + 12 * (64 * 6912 * 768 + 64 * 768 * 12288) * 2 + 256 * (768 * 256 + 2304 * 256) * 2
    return float(llama + n_clip * (clip + proj) + sam)


def run_ours(args, rank, world, dev):
    import ctypes
    from medplib_b200 import _lib
    lib = _lib.load()
    torch.cuda.set_device(dev)
    m = build_model(dev, small=args.small)
    images_clip, images, ids = make_inputs()
    label = torch.zeros(336, 336)
    forced = {SEG_AT: SEG}
    # value: inputs already resident in HBM
    d_clip, d_img, d_ids = images_clip.to(dev).to(bf16), images.to(dev).to(bf16), ids.to(dev)
    # e2e: pinned host buffers, fp32 like the reference's collator output; cast on the device
    h_clip, h_img, h_ids = images_clip.pin_memory(), images.pin_memory(), ids.pin_memory()

    def step_resident():
        return m.evaluate(d_clip, d_img, d_ids, [(256, 256)], [label], max_new_tokens=N_NEW, forced_tokens=forced)

    def step_e2e():
        c = h_clip.to(dev, non_blocking=True).to(bf16)
        i = h_img.to(dev, non_blocking=True).to(bf16)
        t = h_ids.to(dev, non_blocking=True)
        out_ids, masks = m.evaluate(c, i, t, [(256, 256)], [label], max_new_tokens=N_NEW, forced_tokens=forced)
        return out_ids.cpu(), masks[0].cpu()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.mpl_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.mpl_launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    if args.ncu:
        for _ in range(2):
            step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    ms, launches = timed(step_resident, args.steps)
    ck = clocks.stop()
    ms_e2e, _ = timed(step_e2e, args.steps)
    # in-situ durations, CUDA events on the launch stream (the vision / LLM stream overlap is switched off for these
    # passes so every interval is one kernel alone): (1) the persistent decode-step kernel -- the kernel with the largest
    # share of a step -- and (2) every tcgen05 GEMM launch
    m.overlap_vision = False
    psteps = min(args.steps, 3)
    lib.mpl_profile_decode(1)
    for _ in range(psteps):
        step_resident()
    dtot, dcnt = ctypes.c_float(0), ctypes.c_int(0)
    lib.mpl_profile_decode_read(ctypes.byref(dtot), ctypes.byref(dcnt))
    lib.mpl_profile_decode(0)
    lib.mpl_profile_gemm(1)
    for _ in range(psteps):
        step_resident()
    tot, cnt = ctypes.c_float(0), ctypes.c_int(0)
    lib.mpl_profile_gemm_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.mpl_profile_gemm(0)
    m.overlap_vision = True
    pk = peaks()
    gemm_ms = tot.value / psteps
    ach = gemm_flops_per_image() / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 and not args.small else None
    d = DIMS
    T = N_TEXT - 1 + 576
    dk_ms = dtot.value / max(dcnt.value, 1)
    # algorithmic bytes of ONE launch: the active weights once (B = 1: one expert per layer) + the KV cache
    # of the mean context of the step's decode launches; lm_head runs outside the kernel and is not counted
    dk_bytes = d["L"] * (4 * d["D"] ** 2 + 3 * d["D"] * d["F"]) * 2 + 2 * d["L"] * d["D"] * 2 * (T + N_NEW / 2.0)
    dk_ach = dk_bytes / (dk_ms * 1e-3) / 1e9 if dk_ms > 0 and not args.small else None
    if rank != 0:
        return None
    h2d = (h_clip.numel() + h_img.numel()) * 4 + h_ids.numel() * 8
    d2h = 336 * 336 * 2 + (N_TEXT + N_NEW) * 8
    step_ms = ms / args.steps
    line = {
        "metric": "pixel-grounding images/sec at 7B (MedPLIB-7B-2e, bf16, batch 1)", "value": world * args.steps / (ms * 1e-3),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": grounding_config(world, args.small),
        "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": ck,
        # the dominant kernel = the one with the largest share of a step's device time
        "roofline": {"bound": "hbm", "achieved": dk_ach, "peak": pk["hbm"], "unit": "GB/s",
                     "frac": (dk_ach / pk["hbm"]) if dk_ach else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, ncu --set full capture of this kernel
                     # at B = 1, context 615 (profiles/r02_ncu_decode_kernel.md): 13.42 GB read + 0.10 GB written
                     "traffic": None if args.small else 13.52e9, "traffic_source": "profiles/r02_ncu_decode_kernel.md",
                     "kernel": "llama_decode_kernel (one persistent cooperative launch per generated token: all 32 "
                               "layers; algorithmic bytes = one expert's weights + attention weights + KV cache, once)",
                     "algorithmic_bytes_per_launch": dk_bytes, "kernel_ms_per_launch": dk_ms,
                     "kernel_launches_per_step": dcnt.value / psteps,
                     "share_of_step": (dtot.value / psteps) / step_ms if step_ms > 0 else None,
                     "peak_source": pk["src"] + " copy bandwidth"},
        "roofline_gemm": {"bound": "tensor", "achieved": ach, "peak": pk["tf_sus"], "unit": "TFLOP/s",
                          "frac": (ach / pk["tf_sus"]) if ach else None,
                          "traffic": None if args.small else 186.8e6, "traffic_source": "profiles/r02_gemm_tile_widths.md "
                          "(mean DRAM bytes per launch of the 4 LLaMA-layer launches: o_proj, q,k,v, grouped gate|up, "
                          "grouped down)",
                          "kernel": "gemm_bf16_tcgen05_kernel / gemm_bf16_tcgen05_pair_kernel (all launches of a step: "
                                    "algorithmic 2MNK / summed CUDA-event durations)",
                          "kernel_ms_per_step": gemm_ms, "kernel_launches_per_step": cnt.value / psteps,
                          "share_of_step": gemm_ms / step_ms if step_ms > 0 else None,
                          "peak_source": pk["src"] + " sustained bf16"},
    }
    if args.cpu_baseline and world >= 1:
        line["cpu_baseline"] = cpu_reference()
    return line


def grounding_config(world, small=False):
    return {"workload": "MedPLIB-7B pixel-grounding (--eval_seg) bf16, batch 1: evaluate() = CLIP-L/14-336 + "
                        "projector + LLaMA-7B-MoE(2 experts, top-1) prefill T=615 + 8 decode tokens (<SEG> forced)"
                        " + text_hidden_fcs + SAM-Med2D ViT-B@256 + mask decoder + resize 336x336",
            "weights": "random init, 11.07 B params", "parallelism": f"replicas x{world}",
            "l2": "every step streams ~22 GB of weights (>> 126 MB L2), no explicit flush needed",
            "small": bool(small)}


# ------------------------------------------------------------------------------------------------- input pipeline (§8 f-1)
PRE_B, PRE_HW, PRE_ROT = 8, (1024, 1024), 8


def preprocess_batches(n_batches, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [[torch.randint(0, 256, (*PRE_HW, 3), dtype=torch.uint8, generator=g) for _ in range(PRE_B)]
            for _ in range(n_batches)]


def cpu_preprocess(images):
    """The reference's per-image work (ResizeLongestSide via PIL, normalise, pad, CLIP processor) restated by the oracle
    port, one process; also PIL itself (the reference's real dependency) when importable, for scale."""
    import numpy as np
    from oracle import preprocess as op
    t = time.time()
    for im in images:
        op.image_sam(im)
        op.image_clip(im)
    sec = (time.time() - t) / len(images)
    out = {"value": 1.0 / sec, "unit": "images/s", "cores": 1, "kind": "port",
           "sample": f"oracle numpy port of PIL BILINEAR resize (256 + 336 targets) + normalise + pad on {len(images)} "
                     f"synthetic {PRE_HW[0]}x{PRE_HW[1]} images, one process"}
    try:
        from PIL import Image
        lut_s, lut_c = op.sam_lut(), op.clip_lut()
        t = time.time()
        for im in images:
            for L, lut in ((256, lut_s), (336, lut_c)):
                nh, nw = op.get_preprocess_shape(im.shape[0], im.shape[1], L)
                r = np.array(Image.fromarray(im).resize((nw, nh), Image.BILINEAR))
                np.stack([lut[c][r[..., c]] for c in range(3)])
        out["pil_images_per_s_one_core"] = len(images) / (time.time() - t)
    except ImportError:
        pass
    return out


def run_preprocess(args, rank, world, dev):
    """Secondary workload: a step = one batch of 8 decoded 1024x1024 u8 RGB images -> `images`
    [8,3,256,256] + `images_clip` [8,3,336,336] fp32 (the collator's contract) in ONE kernel launch."""
    from medplib_b200 import _lib
    from medplib_b200.preprocess import ImagePreprocessor
    lib = _lib.load()
    torch.cuda.set_device(dev)
    pre = ImagePreprocessor(dev)
    host = preprocess_batches(PRE_ROT, seed=rank)  # 8 batches x 25 MB of inputs + 17 MB of outputs each: > 126 MB L2
    resident = [[im.to(dev) for im in b] for b in host]
    plans = [pre.plan(b) for b in resident]
    state = {"i": 0}

    def step_resident():
        state["i"] += 1
        return pre(resident[state["i"] % PRE_ROT])

    def step_e2e():
        state["i"] += 1
        return pre(host[state["i"] % PRE_ROT])

    def step_kernel():
        state["i"] += 1
        return pre.launch(plans[state["i"] % PRE_ROT])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.mpl_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.mpl_launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    if args.ncu:
        for _ in range(3):
            step_kernel()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_kernel()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
        step_kernel()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    ms, launches = timed(step_resident, args.steps)
    ms_k, _ = timed(step_kernel, args.steps)  # the kernel alone: launches back to back on the stream, events around them
    # the timed regions last a few ms, below nvidia-smi's 100 ms sampling period: keep the same launches running
    # (untimed) until the sampler has seen the clocks under this load
    t_end = time.time() + 0.7
    while len(clocks.rows) < 4 and time.time() < t_end:
        for _ in range(50):
            step_kernel()
        torch.cuda.synchronize()
    ck = clocks.stop()
    ms_e2e, _ = timed(step_e2e, args.steps)
    if rank != 0:
        return
    pk = peaks()
    alg = plans[0]["algorithmic_bytes"]
    ach = alg / (ms_k / args.steps * 1e-3) / 1e9
    line = {
        "metric": "input-pipeline images/sec (decoded u8 -> SAM 256 + CLIP 336 tensors)",
        "value": world * PRE_B * args.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"image input pipeline (LazySupervisedDataset.__getitem__ image branch + collator stack): "
                               f"batch {PRE_B} x {PRE_HW[0]}x{PRE_HW[1]} u8 RGB -> images fp32 [B,3,256,256] + images_clip "
                               f"fp32 [B,3,336,336], PIL-bit-exact", "parallelism": f"replicas x{world}",
                   "l2": f"rotates {PRE_ROT} input/output sets (~340 MB) so no step finds its data in the 126 MB L2",
                   "e2e_note": "outputs stay on the device as the model's input_dict; nothing is read back"},
        "e2e": {"value": world * PRE_B * args.steps / (ms_e2e * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": PRE_B * PRE_HW[0] * PRE_HW[1] * 3 + 2 * PRE_B * 120, "d2h_bytes_per_step": 0},
        "gpu_launches": int(launches), "clocks": ck,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                     # dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full capture of this workload
                     # (first kernel version; the source is read once, the 17 MB of outputs were still in L2 at kernel end)
                     "traffic": 25.23e6, "traffic_source": "profiles/r01_ncu_preprocess.md",
                     "kernel": "preprocess_kernel (one launch per batch)",
                     "algorithmic_bytes_per_launch": alg, "kernel_ms_per_launch": ms_k / args.steps,
                     "peak_source": pk["src"] + " HBM copy bandwidth",
                     "note": "algorithmic bytes = each source image once per target (2x) + every output once"},
    }
    if args.cpu_baseline:
        line["cpu_baseline"] = cpu_preprocess([im.numpy() for im in host[0][:4]])
    return line


def run_reference_preprocess(args, rank, world):
    """--impl reference --workload preprocess: the oracle port on every host core (one image per process)."""
    if rank != 0:
        return
    import concurrent.futures as cf
    cores = os.cpu_count() or 1
    images = [im.numpy() for im in preprocess_batches(1)[0]]
    n = max(PRE_B, cores)
    work = [images[i % PRE_B] for i in range(n)]
    with cf.ProcessPoolExecutor(cores) as ex:
        list(ex.map(_pre_one, work[:cores]))  # warm the workers
        times = []
        for _ in range(max(1, min(args.steps, 3))):
            t = time.time()
            list(ex.map(_pre_one, work))
            times.append(time.time() - t)
    sec = sorted(times)[len(times) // 2]
    val = n / sec
    cb = {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
          "sample": f"oracle numpy port, {n} synthetic {PRE_HW[0]}x{PRE_HW[1]} images per step over {cores} processes"}
    print(json.dumps({
        "impl": "reference", "metric": "input-pipeline images/sec (decoded u8 -> SAM 256 + CLIP 336 tensors)",
        "value": val, "unit": "images/s", "n_gpus": world, "steps": len(times), "warmup": 1,
        "ms_per_step": sec * 1e3 * PRE_B / n, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": {"workload": "same images as the default arm, oracle port on host cores"},
        "cpu_baseline": cb, "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
        flush=True)


def _pre_one(im):
    from oracle import preprocess as op
    op.image_sam(im)
    op.image_clip(im)
    return 0


# ------------------------------------------------------------------------------------------------- VQA decode (configs[2])
def run_decode(args, rank, world, dev):
    """Secondary workload: MedPLIB-7B VQA autoregressive decode, bf16, batch B (default 8), prompt T=615 (40 text ids
    + one 576-token image), `--new-tokens` greedy tokens with EOS disabled. HBM-bound: the roofline is the algorithmic
    bytes of a decode step (active weights once + the KV cache of every sequence) / measured copy bandwidth."""
    from medplib_b200 import _lib
    lib = _lib.load()
    torch.cuda.set_device(dev)
    m = build_model(dev, small=args.small)
    B, new = args.batch, args.new_tokens
    # B different prompts of the same length (different images and token ids, so the sequences route independently)
    _, _, ids0 = make_inputs()
    images_clip = torch.randn(B, 3, 336, 336, generator=torch.Generator().manual_seed(10))
    ids = torch.randint(3, 31999, (B, N_TEXT), generator=torch.Generator().manual_seed(12))
    ids[0] = ids0[0]
    ids[:, 2], ids[:, 3], ids[:, 4] = 32001, -200, 32002
    d_clip = images_clip.to(dev).to(bf16)
    d_ids = ids.to(dev)
    h_clip, h_ids = images_clip.pin_memory(), ids.pin_memory()

    def gen(n, e2e=False):
        if e2e:
            c, t = h_clip.to(dev, non_blocking=True).to(bf16), h_ids.to(dev, non_blocking=True)
            return m.generate(input_ids=t, images=c, max_new_tokens=n, do_sample=False).cpu()
        return m.generate(input_ids=d_ids, images=d_clip, max_new_tokens=n, do_sample=False)

    def timed(fn, reps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.mpl_launch_count()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms / reps, (lib.mpl_launch_count() - n0) // reps

    for _ in range(max(args.warmup, 3)):
        gen(8)
    if args.ncu:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        gen(4)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    gen(8, e2e=True)
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    ms_full, launches = timed(lambda: gen(new), args.steps)
    ck = clocks.stop()
    ms_pre, _ = timed(lambda: gen(1), max(args.steps, 3))
    ms_e2e, _ = timed(lambda: gen(new, e2e=True), args.steps)
    if rank != 0:
        return None
    d = DIMS if not args.small else dict(D=512, F=1024, L=2, H=4, V=32267, E=2)
    step_ms = (ms_full - ms_pre) / max(new - 1, 1)
    T = N_TEXT - 1 + 576
    experts_hit = min(B, d["E"])
    w_bytes = d["L"] * (4 * d["D"] ** 2 + experts_hit * 3 * d["D"] * d["F"]) * 2 + d["V"] * d["D"] * 2
    kv_bytes = 2 * d["L"] * d["D"] * 2 * (T + new / 2.0) * B
    pk = peaks()
    ach = (w_bytes + kv_bytes) / (step_ms * 1e-3) / 1e9
    return {
        "metric": "VQA-decode tokens/sec at 7B (MedPLIB-7B-2e, bf16)", "value": world * B * new / (ms_full * 1e-3),
        "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_full,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"MedPLIB-7B VQA autoregressive decode bf16, batch {B}, prompt T={T} (40 ids + 576 image "
                               f"tokens), {new} new tokens greedy, EOS disabled; generate() incl. CLIP + prefill",
                   "weights": "random init, 11.07 B params", "parallelism": f"replicas x{world}",
                   "l2": "every decode step streams >= 13 GB of weights (>> 126 MB L2)", "small": bool(args.small)},
        "e2e": {"value": world * B * new / (ms_e2e * 1e-3), "unit": "tokens/s",
                "h2d_bytes_per_step": h_clip.numel() * 4 + h_ids.numel() * 8, "d2h_bytes_per_step": B * (N_TEXT + new) * 8},
        "gpu_launches": int(launches), "clocks": ck,
        "decode_step_ms": step_ms, "prefill_ms": ms_pre, "decode_tokens_per_s": world * B / (step_ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                     "traffic": None, "kernel": "decode step (streaming GEMMs + KV-cache attention): algorithmic bytes = "
                     f"active weights {w_bytes / 1e9:.2f} GB + mean KV {kv_bytes / 1e9:.2f} GB per step / step time",
                     "peak_source": pk["src"] + " copy bandwidth"}}


# ------------------------------------------------------------------------------------------------- train step (configs[3])
N_TEXT_TRAIN = 64


def train_batch(B, seed):
    g = torch.Generator().manual_seed(100 + seed)
    ids = torch.randint(3, 31999, (B, N_TEXT_TRAIN), generator=g)
    ids[:, 2], ids[:, 3], ids[:, 4] = 32001, -200, 32002
    ids[:, 40] = SEG
    labels = ids.clone()
    labels[:, :24] = -100  # the prompt part carries no loss (LazySupervisedDataset masks the human turn)
    am = torch.ones_like(ids, dtype=torch.bool)
    images_clip = torch.randn(B, 3, 336, 336, generator=g)
    images = torch.randn(B, 3, 256, 256, generator=g)
    gts = [(torch.rand(336, 336, generator=g) > 0.7).float() for _ in range(B)]
    # a region slot (-300) with a 24x24 blob mask in every second sample (rp_flag batches, medplib_arch.py:426-429)
    region_masks, valid = [], []
    for b in range(B):
        if b % 2 == 0:
            ids[b, 12] = -300
            labels[b, 12] = -100
            m = torch.zeros(24, 24)
            y0, x0 = int(torch.randint(0, 12, (1,), generator=g)), int(torch.randint(0, 12, (1,), generator=g))
            m[y0:y0 + 10, x0:x0 + 10] = 1.0
            region_masks.append([m])
            valid.append([True])
        else:
            valid.append([False])
    return ids, labels, am, images_clip, images, gts, region_masks, valid


def run_train(args, rank, world, dev):
    """Secondary workload: Stage-IV-flags train step (scripts/train_stage4.sh: --moe_enable, LoRA r=8 alpha=16 on
    q,v,gate,up,down, sft wg,lm_head,embed_tokens,mask_decoder,text_hidden_fcs,region_fea_adapter; capacity 1.5, aux
    coef 0), bf16, micro-batch `--batch` per GPU (default 8), T = 64 text ids + 575 = 639, one <SEG> + one 336x336
    ground-truth mask per sample. A step = forward (activations kept) + hand-scheduled backward + bucketed NCCL
    all-reduce of the fp32 gradient arena (N > 1) + clip + AdamW. Data-parallel: `scaling` weak, value = samples of all
    ranks / max-over-ranks time."""
    import ctypes
    from medplib_b200 import _lib, train
    lib = _lib.load()
    torch.cuda.set_device(dev)
    m = build_model(dev, small=args.small)
    m.config.moe["router_aux_loss_coef"] = 0.0
    m.router_aux_loss_coef = 0.0
    m.ce_loss_weight, m.bce_loss_weight, m.dice_loss_weight, m.iou_loss_weight, m.focal_loss_weight = 1.0, 2.0, 0.5, 1.0, 1.0
    train.attach_lora(m, r=8, lora_alpha=16, lora_dropout=0.0, target_modules="q_proj,v_proj,gate_proj,up_proj,down_proj")
    train.set_trainable(m, "wg,lm_head,embed_tokens,mask_decoder,text_hidden_fcs,region_fea_adapter")
    m.train()
    tr = m.trainer(lr=3e-4)
    B = args.batch
    ids, labels, am, images_clip, images, gts, region_masks, valid = train_batch(B, rank)
    d_rm = [[x.to(dev) for x in r] for r in region_masks]
    d_in = dict(ids=ids.to(dev), labels=labels.to(dev), am=am.to(dev), clip=images_clip.to(dev).to(bf16),
                img=images.to(dev).to(bf16), gts=[x.to(dev) for x in gts])
    h_in = dict(ids=ids.pin_memory(), labels=labels.pin_memory(), am=am.pin_memory(), clip=images_clip.pin_memory(),
                img=images.pin_memory(), gts=[x.pin_memory() for x in gts])

    def step(x, read_loss=False):
        out = m(images=x["img"], images_clip=x["clip"], input_ids=x["ids"], region_masks=d_rm,
                valid_region_masks_bool=valid, labels=x["labels"], attention_mask=x["am"], offset=None,
                masks_list=x["gts"], label_list=x["gts"], resize_list=[(256, 256)] * B, inference=False, seg_flag=True,
                rp_flag=True)
        out["loss"].backward()
        tr.step()
        return float(out["loss"].detach()) if read_loss else None

    def step_e2e():
        x = dict(ids=h_in["ids"].to(dev, non_blocking=True), labels=h_in["labels"].to(dev, non_blocking=True),
                 am=h_in["am"].to(dev, non_blocking=True), clip=h_in["clip"].to(dev, non_blocking=True).to(bf16),
                 img=h_in["img"].to(dev, non_blocking=True).to(bf16),
                 gts=[t.to(dev, non_blocking=True) for t in h_in["gts"]])
        return step(x, read_loss=True)

    def timed(fn, steps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.mpl_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, lib.mpl_launch_count() - n0

    if args.ncu:
        for _ in range(2):
            step(d_in)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(d_in)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    loss0 = None
    for i in range(max(args.warmup, 3)):
        l = step(d_in, read_loss=True)
        loss0 = l if loss0 is None else loss0
    step_e2e()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    ms, launches = timed(lambda: step(d_in), args.steps)
    ck = clocks.stop()
    ms_e2e, _ = timed(step_e2e, args.steps)
    loss1 = step(d_in, read_loss=True)
    lib.mpl_profile_gemm(1)
    psteps = min(args.steps, 2)
    for _ in range(psteps):
        step(d_in)
    tot, cnt = ctypes.c_float(0), ctypes.c_int(0)
    lib.mpl_profile_gemm_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.mpl_profile_gemm(0)
    if rank != 0:
        return None
    d = DIMS if not args.small else dict(D=512, F=1024, L=2, H=4, V=32267, E=2)
    T = N_TEXT_TRAIN - 1 + 576
    lin = T * d["L"] * (4 * d["D"] ** 2 + 3 * d["D"] * d["F"]) * 2
    head = T * d["D"] * d["V"] * 2
    enc = gemm_flops_per_image(T) - (N_TEXT - 1 + 576) * DIMS["L"] * (4 * DIMS["D"] ** 2 + 3 * DIMS["D"] * DIMS["F"]) * 2
    flops = B * (2 * lin + 3 * head + enc)  # fwd + dgrad through frozen weights; lm_head adds its wgrad; encoders fwd
    pk = peaks()
    gemm_ms = tot.value / psteps
    ach = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 and not args.small else None
    h2d = sum(t.numel() * t.element_size() for t in (ids, labels, am, images_clip, images)) + sum(g.numel() * 4 for g in gts)
    return {
        "metric": "Stage-IV train step samples/sec at 7B (MedPLIB-7B-2e, bf16, LoRA r=8 + sft modules)",
        "value": world * B * args.steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"Stage-IV-flags train step bf16, micro-batch {B}/GPU x {world} GPU (global {B * world}), "
                               f"T={T} (64 ids + 576 image tokens), 1 <SEG> + 336x336 GT mask per sample; MoE dense 32 "
                               "layers E=2 top-1 cf=1.5; LoRA r=8 a=16 on q,v,gate,up,down; sft wg,lm_head,embed_tokens,"
                               "mask_decoder,text_hidden_fcs,region_fea_adapter; a 24x24 region slot in every second "
                               "sample; fwd + bwd + grad all-reduce + clip + AdamW",
                   "weights": "random init, 11.07 B params", "parallelism": f"dp{world}",
                   "trainable_params": int(tr.arena.numel), "l2": "activations + weights >> 126 MB L2",
                   "small": bool(args.small)},
        "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches // max(args.steps, 1)), "clocks": ck, "loss_first": loss0, "loss_last": loss1,
        "tokens_per_s": world * B * T * args.steps / (ms * 1e-3),
        "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["tf_sus"], "unit": "TFLOP/s",
                     "frac": (ach / pk["tf_sus"]) if ach else None, "traffic": None,
                     "kernel": "gemm_bf16_tcgen05_kernel (all launches of a train step: forward + dgrad + lm_head wgrad;"
                               " algorithmic 2MNK / summed CUDA-event durations)", "kernel_ms_per_step": gemm_ms,
                     "kernel_launches_per_step": cnt.value / psteps, "step_tflops": flops / (ms / args.steps * 1e-3) / 1e12,
                     "peak_source": pk["src"] + " sustained bf16"}}


# ------------------------------------------------------------------------------------------------- MedPLIB-ICL (configs[4])
ICL_N_IMG, ICL_N_MASK, ICL_N_TEXT = 4, 3, 90


def icl_inputs(seed=0):
    """3 (image, mask) exemplars + the query image in separate mode, every image compressed
    576 -> 256 tokens, every exemplar mask encoded to 64 tokens, ~90 text ids with 7 IMAGE sentinels and <SEG> in the
    prompt: T = 90 - 7 + 4 * 256 + 3 * 64 = 1299."""
    g = torch.Generator().manual_seed(40 + seed)
    ids = torch.randint(3, 31999, (1, ICL_N_TEXT), generator=g)
    types_ = [["image", "mask"] * ICL_N_MASK + ["image"]]
    lengths = [[256, 64] * ICL_N_MASK + [256]]
    for k in range(ICL_N_IMG + ICL_N_MASK):
        ids[0, 4 + 8 * k] = -200
    ids[0, ICL_N_TEXT - 6] = SEG
    clip = torch.randn(ICL_N_IMG, 3, 336, 336, generator=g)
    masks = (torch.rand(ICL_N_MASK, 1, 336, 336, generator=g) < 0.2).float()
    sam = torch.randn(1, 3, 256, 256, generator=g)
    return ids, clip, masks, sam, types_, lengths


def run_icl(args, rank, world, dev):
    """Secondary workload (BASELINE configs[4]): MedPLIB-ICL separate mode, single-pass model_forward(inference=True)
    = 4x CLIP-L/14-336 -> projector -> TokenCompressor (pool 576 -> 256 + LayerNorm + Linear) ; MaskTokenEncoder on 3
    exemplar masks ; splice (T = 1299) ; LLaMA-7B-MoE prefill ; [SEG] row -> text_hidden_fcs -> SAM-Med2D -> mask.
    A step = one query image with its 3 exemplars. Tensor-core bound (prefill); the compressor's pool + LayerNorm kernel
    is the HBM-bound piece north_star names: its in-situ duration and GB/s are reported separately."""
    import ctypes
    from medplib_b200 import _lib, ops
    lib = _lib.load()
    torch.cuda.set_device(dev)
    m = build_model(dev, small=args.small, icl=True)
    ids, clip, masks, sam, types_, lengths = icl_inputs(rank)
    label = torch.zeros(336, 336)
    d = dict(ids=ids.to(dev), clip=clip.to(dev).to(bf16), masks=masks.to(dev).to(bf16), sam=sam.to(dev).to(bf16))
    h = dict(ids=ids.pin_memory(), clip=clip.pin_memory(), masks=masks.pin_memory(), sam=sam.pin_memory())
    am = torch.ones_like(ids, dtype=torch.bool).to(dev)

    def fwd(x):
        return m(images=x["sam"], images_clip=[x["clip"]], input_ids=x["ids"], region_masks=None, labels=None,
                 attention_mask=am, offset=None, masks_list=[label], label_list=[label], resize_list=[(256, 256)],
                 inference=True, mask_images=[x["masks"]], image_token_types=types_, image_token_lengths=lengths,
                 icl_image_counts=[ICL_N_IMG])

    def step_resident():
        return fwd(d)

    def step_e2e():
        x = dict(ids=h["ids"].to(dev, non_blocking=True), clip=h["clip"].to(dev, non_blocking=True).to(bf16),
                 masks=h["masks"].to(dev, non_blocking=True).to(bf16), sam=h["sam"].to(dev, non_blocking=True).to(bf16))
        return fwd(x)["pred_masks"][0].cpu()

    def timed(fn, steps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.mpl_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, lib.mpl_launch_count() - n0

    if args.ncu:
        for _ in range(2):
            step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    for _ in range(max(args.warmup, 3)):
        out = step_resident()
        step_e2e()
    T = ICL_N_TEXT - (ICL_N_IMG + ICL_N_MASK) + ICL_N_IMG * 256 + ICL_N_MASK * 64
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    ms, launches = timed(step_resident, args.steps)
    ck = clocks.stop()
    ms_e2e, _ = timed(step_e2e, args.steps)
    m.overlap_vision = False
    lib.mpl_profile_gemm(1)
    psteps = min(args.steps, 3)
    for _ in range(psteps):
        step_resident()
    tot, cnt = ctypes.c_float(0), ctypes.c_int(0)
    lib.mpl_profile_gemm_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.mpl_profile_gemm(0)
    m.overlap_vision = True
    # the compressor's pool + LayerNorm kernel alone (HBM bound): [4, 576, 4096] bf16 -> [4, 256, 4096]; rotate over 8
    # input buffers (8 x 18.9 MB in + 8.4 MB out > 126 MB L2)
    comp = m.get_model().mm_token_compressor
    xs = [torch.randn(ICL_N_IMG, 576, DIMS["D"] if not args.small else 512, device=dev).to(bf16) for _ in range(8)]
    for x in xs[:3]:
        ops.pool_layernorm(x, comp.norm.weight, comp.norm.bias, 256, comp.norm.eps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(32):
        ops.pool_layernorm(xs[i % 8], comp.norm.weight, comp.norm.bias, 256, comp.norm.eps)
    e1.record()
    torch.cuda.synchronize()
    pool_ms = e0.elapsed_time(e1) / 32
    if rank != 0:
        return None
    dd = DIMS if not args.small else dict(D=512, F=1024, L=2, H=4, V=32267, E=2)
    pool_bytes = ICL_N_IMG * (576 + 256) * dd["D"] * 2
    pk = peaks()
    # decoder linears over T positions + 4 x (CLIP + projector) + SAM encoder + the compressor's Linear on 4 x 256 tokens
    flops = gemm_flops_per_image(T, n_clip=ICL_N_IMG) + ICL_N_IMG * 256 * dd["D"] * dd["D"] * 2
    gemm_ms = tot.value / psteps
    ach = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 and not args.small else None
    step_ms = ms / args.steps
    return {
        "metric": "MedPLIB-ICL pixel-grounding images/sec at 7B (separate mode, 3 exemplars, 576->256 compression)",
        "value": world * args.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"MedPLIB-ICL separate-mode, 3 (img,mask) exemplars + query, mm_token_compress 576->256, "
                               f"mask encoder 64 tokens, T={T}, <SEG> in the prompt, model_forward(inference=True), bf16, "
                               "batch 1", "weights": "random init", "parallelism": f"replicas x{world}",
                   "l2": "every step streams > 13 GB of weights (>> 126 MB L2)", "small": bool(args.small)},
        "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": (clip.numel() + masks.numel() + sam.numel()) * 4 + ids.numel() * 8,
                "d2h_bytes_per_step": 336 * 336 * 2},
        "gpu_launches": int(launches // max(args.steps, 1)), "clocks": ck,
        "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["tf_sus"], "unit": "TFLOP/s",
                     "frac": (ach / pk["tf_sus"]) if ach else None, "traffic": None,
                     "kernel": "gemm_bf16_tcgen05_kernel (all launches of a step; algorithmic 2MNK / summed CUDA-event "
                               "durations)", "kernel_ms_per_step": gemm_ms, "kernel_launches_per_step": cnt.value / psteps,
                     "share_of_step": gemm_ms / step_ms if step_ms > 0 else None,
                     "peak_source": pk["src"] + " sustained bf16"},
        "roofline_pool_ln": {"bound": "hbm", "achieved": pool_bytes / (pool_ms * 1e-3) / 1e9, "peak": pk["hbm"],
                             "unit": "GB/s", "frac": pool_bytes / (pool_ms * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                             "kernel": "pool_ln_kernel (TokenCompressor: 3-token windowed mean 576 -> 256 + LayerNorm, "
                                       f"N = {ICL_N_IMG} images): bytes in + out once",
                             "algorithmic_bytes_per_launch": pool_bytes, "kernel_ms_per_launch": pool_ms,
                             "peak_source": pk["src"] + " copy bandwidth"}}


# ------------------------------------------------------------------------------------------------- reference arm (CPU)
_CPU_STATE = {}


def _cpu_state():
    """Weights for the CPU arm, built once: the full CLIP-L tower (24 layers) and SAM-Med2D ViT-B encoder, the mask
    head, and ONE LLaMA-MoE layer at 7B width whose tensors are ALIASED as all 32 layers (a distinct copy per layer
    would be 44 GB of fp32 and two minutes of random-number generation; a layer's 1.35 GB is far beyond any CPU cache,
    so the time of a full-depth pass is the same)."""
    if _CPU_STATE:
        return _CPU_STATE
    from oracle import weights
    d = DIMS
    moe = dict(num_experts=2, top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0, min_capacity=0,
               router_aux_loss_coef=0.01)
    one = dict(hidden_size=d["D"], intermediate_size=d["F"], num_layers=1, num_heads=d["H"], vocab_size=64,
               rms_norm_eps=1e-5, max_position_embeddings=4096, rope_theta=1e4, moe=moe)
    sd1 = weights.llama(one, seed=0, dtype=torch.float32)
    sd = dict(sd1)
    for k, v in sd1.items():
        if k.startswith("model.layers.0."):
            for l in range(1, d["L"]):
                sd["model.layers.%d.%s" % (l, k[len("model.layers.0."):])] = v
    lcfg = dict(one, num_layers=d["L"])
    ccfg = dict(hidden_size=1024, intermediate_size=4096, num_layers=24, num_heads=16, image_size=336, patch_size=14)
    scfg = dict(embed_dim=768, depth=12, num_heads=12, image_size=256, patch_size=16, out_chans=256)
    g = torch.Generator().manual_seed(5)
    _CPU_STATE.update(lcfg=lcfg, sd=sd, ccfg=ccfg, csd=weights.clip(ccfg, seed=1, dtype=torch.float32), scfg=scfg,
                      ssd=weights.sam_encoder(scfg, seed=2, dtype=torch.float32),
                      hsd=weights.sam_head(seed=3, dtype=torch.float32),
                      proj=[torch.randn(d["D"], 1024, generator=g) * 0.02, torch.randn(d["D"], d["D"], generator=g) * 0.02],
                      fcs=[torch.randn(d["D"], d["D"], generator=g) * 0.02, torch.randn(256, d["D"], generator=g) * 0.02],
                      lm_head=torch.randn(d["V"], d["D"], generator=g) * 0.02)
    return _CPU_STATE


def cpu_image(depth_scale=1.0):
    """ONE image of the benchmark's workload through the reference-semantics oracle port on the host cores, fp32, all
    threads, FULL depth: CLIP-L (23 of 24 layers, as hidden_states[-2]) -> projector -> LLaMA-MoE 32 layers prefill
    T = 615 (+ lm_head on the last row) -> 7 KV-cached decode steps (each + lm_head) -> text_hidden_fcs -> SAM-Med2D
    encoder (12 blocks) -> prompt encoder + mask decoder -> resize. Returns seconds. depth_scale < 1 (warm-up only) runs
    proportionally fewer decoder layers."""
    from oracle import clip, llama, sam, heads
    st = _cpu_state()
    d = DIMS
    torch.set_num_threads(os.cpu_count() or 1)
    lcfg = dict(st["lcfg"], num_layers=max(1, int(round(d["L"] * depth_scale))))
    T = N_TEXT - 1 + 576
    img_c, img_s = torch.randn(1, 3, 336, 336), torch.randn(1, 3, 256, 256)
    t0 = time.time()
    with torch.no_grad():
        feats = clip.vision_tower(st["csd"], "", img_c, st["ccfg"], select_layer=-2)
        img_tok = torch.nn.functional.gelu(feats @ st["proj"][0].T) @ st["proj"][1].T
        x = torch.cat([torch.randn(1, T - 576, d["D"]), img_tok], 1)
        out = llama.model_forward(st["sd"], lcfg, x)
        kv = out["past_key_values"]
        (out["last_hidden_state"][:, -1] @ st["lm_head"].T).argmax(-1)
        hs = out["last_hidden_state"][:, -1:]
        for s in range(N_NEW - 1):
            o = llama.model_forward(st["sd"], lcfg, torch.randn(1, 1, d["D"]), torch.ones(1, T + s + 1, dtype=torch.bool), kv)
            kv = o["past_key_values"]
            (o["last_hidden_state"][:, -1] @ st["lm_head"].T).argmax(-1)
            hs = o["last_hidden_state"]
        prompt = torch.relu(hs @ st["fcs"][0].T) @ st["fcs"][1].T
        emb = sam.image_encoder(st["ssd"], "", img_s, num_heads=12)
        dpe = sam.dense_pe(st["hsd"], "prompt_encoder.", (16, 16))
        sp, de = sam.prompt_encoder_text(st["hsd"], "prompt_encoder.", prompt.reshape(1, 1, 256), (16, 16))
        low, _ = sam.mask_decoder(st["hsd"], "mask_decoder.", emb, dpe, sp, de)
        heads.postprocess_masks(low, (256, 256), (336, 336))
    return time.time() - t0


def cpu_reference():
    """cpu_baseline of the default arm: ONE full-depth image (about 10 s of CPU work on 16 cores), see cpu_image."""
    t0 = time.time()
    _cpu_state()
    setup = time.time() - t0
    sec = cpu_image()
    return {"value": 1.0 / sec, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "oracle port (oracle/: transformers-4.31 / DeepSpeed-0.13.1 semantics restated), fp32, all host "
                      "threads, ONE image of the same workload at FULL depth and 7B width (CLIP-L 23 layers, projector, 32 "
                      "LLaMA-MoE layers prefill T=615 + 7 KV-cached decode steps + lm_head, text_hidden_fcs, SAM-Med2D "
                      f"ViT-B, mask head); the 32 decoder layers alias one layer's weights; measured {sec:.1f} s "
                      f"(weight setup {setup:.0f} s untimed)"}


def run_reference(args, rank, world):
    """--impl reference: K timed steps, each ONE full image through the CPU port at full depth (cpu_image); W warm-up
    passes at 1/8 depth (they only warm the thread pool and the allocator). Rank 0 alone works."""
    if rank != 0:
        return None
    _cpu_state()
    for _ in range(args.warmup):
        cpu_image(depth_scale=0.125)
    times = []
    t0 = time.time()
    for _ in range(args.steps):
        times.append(cpu_image())
    wall = time.time() - t0
    sec = wall / max(len(times), 1)
    val = 1.0 / sec
    cores = os.cpu_count()
    return {
        "impl": "reference", "metric": "pixel-grounding images/sec at 7B (MedPLIB-7B-2e, bf16, batch 1)", "value": val,
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": grounding_config(world),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "the reference's path on the host cores: oracle port (the reference's third-party "
                                   "arithmetic -- transformers 4.31, deepspeed 0.13.1 -- is not installable offline), fp32,"
                                   f" {cores} threads, every step = one full-depth image of the same workload "
                                   f"(see cpu_image); min {min(times):.1f} s, max {max(times):.1f} s per image"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "grounding", "decode", "train", "icl", "preprocess"],
                    help="all (default): the grounding line with the other BASELINE configs under `secondary`")
    ap.add_argument("--batch", type=int, default=8, help="decode / train workloads: sequences (samples) per GPU")
    ap.add_argument("--new-tokens", type=int, default=512, help="decode workload: generated tokens per sequence")
    ap.add_argument("--small", action="store_true", help="2-layer toy LLaMA (plumbing check, not a benchmark)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="--workload all without the secondary workloads (= --workload grounding)")
    ap.add_argument("--ncu", action="store_true",
                    help="profiling aid: warm up, then run ONE step between cudaProfilerStart/Stop (use with "
                         "ncu --profile-from-start off); prints no bench line")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: Python-level prints keep the real stdout, anything a library writes to
    # file descriptor 1 from C (NCCL's version banner at communicator creation) is sent to stderr instead
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        line = (run_reference_preprocess if args.workload == "preprocess" else run_reference)(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: medplib_b200 has no CPU path (use --impl reference for the CPU arm)")
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL's version banner goes to stdout at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        torch.cuda.set_device(dev)
        torch.distributed.init_process_group("nccl", device_id=dev)
    args.cpu_baseline = args.cpu_baseline and rank == 0 and world == 1
    runners = {"grounding": run_ours, "decode": run_decode, "train": run_train, "icl": run_icl,
               "preprocess": run_preprocess}
    if args.workload != "all":
        line = runners[args.workload](args, rank, world, dev)
    else:
        line = run_ours(args, rank, world, dev)
        if args.secondary and not args.ncu:
            # the other BASELINE configs, same process, same GPUs, one after the other (each frees its model first);
            # short fixed step counts so the default run stays within a few minutes
            import copy
            import gc
            sec = {}
            for name, steps, warm in (("decode", 1, 3), ("train", 5, 3), ("icl", 10, 3)):
                gc.collect()
                torch.cuda.empty_cache()
                a2 = copy.copy(args)
                a2.steps, a2.warmup, a2.cpu_baseline = steps, warm, False
                try:
                    sec[name] = runners[name](a2, rank, world, dev)
                except Exception as e:  # a secondary workload must not take the headline line down with it
                    sec[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                    if world > 1:
                        raise
            if line is not None:
                line["secondary"] = sec
    if line is not None and rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
