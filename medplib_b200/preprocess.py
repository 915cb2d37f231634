"""GPU image input pipeline: the step immediately before the hot path.

``ImagePreprocessor`` replaces, for a batch of decoded RGB images, the image branch of the reference's
``LazySupervisedDataset.__getitem__`` (datasets/LazySupervisedDataset.py:516-519, 535-553: ``ResizeLongestSide`` through
PIL, ``preprocess`` / ``pad_tensor_channelwise``, ``CLIPImageProcessor.preprocess``) plus the stacking the collator does
(datasets/DataCollatorForSupervisedDataset.py:110-138), and returns the same keys ``forward(**input_dict)`` consumes:
``images`` [B,3,256,256], ``images_clip`` [B,3,336,336], ``resize_list``; optionally ``region_masks``.  The u8 images are
copied to the device once (ragged, one staging buffer) and ONE kernel launch (csrc/preprocess.cu) produces every output;
results are bit-identical to the reference's CPU pipeline.  The host only builds PIL's per-axis coefficient tables (a few
KB, cached per (in, out) size) and the 256-entry value tables.  No CPU fallback: without the CUDA library this raises.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib

PRECISION_BITS = 32 - 8 - 2  # Pillow libImaging/Resample.c

# datasets/LazySupervisedDataset.py:394-399
SAM_PIXEL_MEAN = (123.675, 116.28, 103.53)
SAM_PIXEL_STD = (58.395, 57.12, 57.375)
CLIP_IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)


def get_preprocess_shape(oldh, oldw, long_side_length):
    """ResizeLongestSide.get_preprocess_shape (model/segment_anything_med2d/utils/transforms.py:95-103)."""
    scale = long_side_length * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


def pil_coeffs(in_size, out_size):
    """PIL's BILINEAR resampling tables for one axis (precompute_coeffs + normalize_coeffs_8bpc, Resample.c):
    bounds int32 [out, 2] = (first tap, tap count), coeffs int32 [out, ksize] in 22-bit fixed point."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = filterscale  # bilinear support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)  # C (int) cast truncates; operands are > -1
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    w = np.maximum(1.0 - np.abs((x + xmin[:, None] - center[:, None] + 0.5) * (1.0 / filterscale)), 0.0)
    w = np.where(x < xmax[:, None], w, 0.0)
    ww = np.add.accumulate(w, axis=1)[:, -1:]  # sequential sum, like the C loop (trailing zeros do not change it)
    w = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    coeffs = (0.5 + w * (1 << PRECISION_BITS)).astype(np.int32)  # weights are >= 0: the -0.5 branch never fires
    return np.stack([xmin, xmax], 1).astype(np.int32), coeffs


def band_rows_bound(in_size, out_size, R):
    """Mirror of csrc/preprocess.cu:band_rows_bound (integer-only upper bound of the source rows under R output rows)."""
    sup = max((in_size + out_size - 1) // out_size, 1)
    span = ((R - 1) * in_size + out_size - 1) // out_size
    return min(span + 2 * sup + 2, in_size)


def sam_level_table():
    """fp32 [3,256]: (level - pixel_mean) / pixel_std in fp32 (LazySupervisedDataset.py:484)."""
    lv = torch.arange(256, dtype=torch.float32)[None, :]
    return (lv - torch.tensor(SAM_PIXEL_MEAN).view(3, 1)) / torch.tensor(SAM_PIXEL_STD).view(3, 1)


def clip_level_table():
    """fp32 [3,256]: CLIPImageProcessor (transformers 4.31) rescale — u8 * (1/255) in float64, cast to fp32 — then
    normalize (x - mean) / std in fp32."""
    lv = (np.arange(256, dtype=np.uint8) * 0.00392156862745098).astype(np.float32)[None, :]
    mean = np.array(CLIP_IMAGE_MEAN, np.float32)[:, None]
    std = np.array(CLIP_IMAGE_STD, np.float32)[:, None]
    return torch.from_numpy(((lv - mean) / std).astype(np.float32))


def clip_pad_levels():
    """u8 level per channel of the CLIP branch's padding: (mean * 255).clamp(0, 255).to(int) (LazySupervisedDataset.py:398)."""
    return (torch.tensor(CLIP_IMAGE_MEAN) * 255).clamp(0, 255).to(torch.int).tolist()


class PreprocessJob(ctypes.Structure):
    """mpl_preprocess_job (include/medplib_b200.h)."""
    _fields_ = [
        ("src", _lib.c_void_p), ("src_stride", _lib.c_ll), ("H", _lib.c_int), ("W", _lib.c_int), ("C", _lib.c_int),
        ("new_h", _lib.c_int), ("new_w", _lib.c_int), ("out_size", _lib.c_int), ("pad_top", _lib.c_int),
        ("pad_left", _lib.c_int), ("coef_x", _lib.c_void_p), ("bound_x", _lib.c_void_p), ("coef_y", _lib.c_void_p),
        ("bound_y", _lib.c_void_p), ("ks_x", _lib.c_int), ("ks_y", _lib.c_int), ("lut", _lib.c_void_p),
        ("pad_value", _lib.c_float * 3), ("out_dtype", _lib.c_int), ("dst", _lib.c_void_p),
    ]


_DT = {torch.bfloat16: _lib.DT_BF16, torch.float32: _lib.DT_F32, torch.uint8: 2}


class ImagePreprocessor:
    """``pre(images, region_masks=None) -> dict`` with the collator's keys, computed on ``device``.

    images: list of u8 RGB arrays / tensors [H, W, 3] (what ``cv2.cvtColor(cv2.imread(p), cv2.COLOR_BGR2RGB)`` returns),
    any sizes.  region_masks: optional list (per image) of lists of u8 {0,1} [H, W] masks; returned as the 24x24 grids
    the reference derives before its connected-component filter (LazySupervisedDataset.py:516-519).  encoder_masks:
    optional list (per sample) of lists of u8 {0,1} [H, W] ICL exemplar masks (any sizes); returned as `mask_images`,
    the {0,1} [m, 1, 336, 336] inputs of the MaskTokenEncoder (ICLLazySupervisedDataset.py:77-85).
    out_dtype: torch.float32 (the reference's contract, a-0) or torch.bfloat16 (what the model computes in; the same
    values rounded once, i.e. what ``.to(bfloat16)`` of the fp32 tensors gives).
    """

    def __init__(self, device="cuda", sam_size=256, clip_size=336, patch=14, out_dtype=torch.float32):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MplError("medplib_b200.preprocess needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device)
        self.sam_size, self.clip_size, self.patch, self.out_dtype = sam_size, clip_size, patch, out_dtype
        self.sam_lut = sam_level_table().contiguous().to(self.device)
        clip = clip_level_table()
        self.clip_lut = clip.contiguous().to(self.device)
        self.clip_pad = [float(clip[c, lvl]) for c, lvl in enumerate(clip_pad_levels())]
        self.enc_lut = (torch.arange(256) > 0).float().view(1, 256).contiguous().to(self.device)  # "resized mask > 0"
        self._tables = {}

    def _axis(self, n_in, n_out, tap_major):
        """Device tables of one axis. The horizontal pass reads its coefficients tap-major ([ks4, n_out], tap count
        rounded up to a multiple of 4 with zeros: a warp's 32 coefficients of one tap are one 128-byte line and the
        kernel consumes taps four at a time); the vertical pass reads PIL's own [n_out, ks] layout."""
        key = (n_in, n_out, tap_major)
        t = self._tables.get(key)
        if t is None:
            bounds, coeffs = pil_coeffs(n_in, n_out)
            ks = coeffs.shape[1]
            if tap_major:
                ks = (ks + 3) // 4 * 4
                padded = np.zeros((ks, n_out), np.int32)
                padded[:coeffs.shape[1]] = coeffs.T
                coeffs = padded
            t = (torch.from_numpy(bounds).contiguous().to(self.device),
                 torch.from_numpy(np.ascontiguousarray(coeffs)).to(self.device), ks)
            if len(self._tables) > 4096:
                self._tables.clear()
            self._tables[key] = t
        return t

    def _job(self, src, H, W, C, target, lut, pad_value, dst):
        new_h, new_w = get_preprocess_shape(H, W, target)
        if new_h < 1 or new_w < 1:  # the reference's PIL resize raises on an empty target as well
            raise ValueError(f"a {H}x{W} image has no pixels left at longest side {target}")
        bx, cx, ksx = self._axis(W, new_w, True)
        by, cy, ksy = self._axis(H, new_h, False)
        j = PreprocessJob()
        j.src, j.src_stride, j.H, j.W, j.C = src, W * C, H, W, C
        j.new_h, j.new_w, j.out_size = new_h, new_w, target
        j.pad_top, j.pad_left = (target - new_h) // 2, (target - new_w) // 2
        j.coef_x, j.bound_x, j.coef_y, j.bound_y = cx.data_ptr(), bx.data_ptr(), cy.data_ptr(), by.data_ptr()
        j.ks_x, j.ks_y = ksx, ksy
        j.lut = lut.data_ptr() if lut is not None else None
        j.pad_value = (ctypes.c_float * 3)(*pad_value)
        j.out_dtype, j.dst = _DT[dst.dtype], dst.data_ptr()
        return j, (new_h, new_w)

    def plan(self, images, region_masks=None, encoder_masks=None):
        """Validate, stage the u8 inputs on the device (host inputs: one pinned ragged buffer, one copy; device inputs are
        used in place) and build the job list.  Returns a plan ``launch`` runs; ``__call__`` = ``launch(plan(...))``."""
        B = len(images)
        arrays = [torch.as_tensor(im) for im in images]
        masks = [[torch.as_tensor(m) for m in (region_masks[i] if region_masks is not None else [])] for i in range(B)]
        for a in arrays:
            if a.dtype != torch.uint8 or a.dim() != 3 or a.shape[2] != 3:
                raise ValueError("images must be uint8 [H, W, 3] RGB")
        for i, ms in enumerate(masks):
            for m in ms:
                if m.dtype != torch.uint8 or tuple(m.shape) != tuple(arrays[i].shape[:2]):
                    raise ValueError("region masks must be uint8 [H, W] of their image's size")
        # ICL exemplar masks for the MaskTokenEncoder (ICLLazySupervisedDataset._preprocess_encoder_mask :77-85): the
        # reference resizes mask * 255, so the {0,1} masks are scaled while they are staged
        enc = [[torch.as_tensor(m) for m in ms] for ms in (encoder_masks or [])]
        for ms in enc:
            for m in ms:
                if m.dtype != torch.uint8 or m.dim() != 2:
                    raise ValueError("encoder masks must be uint8 {0,1} [H, W]")
        enc_flat = [(m != 0).to(torch.uint8) * 255 for ms in enc for m in ms]
        flat = [a for a in arrays] + [m for ms in masks for m in ms] + enc_flat
        keep, ptrs = [], []
        if flat and all(f.device.type == "cuda" for f in flat):
            for f in flat:
                f = f.contiguous()
                if f.data_ptr() % 16:  # the kernel stages aligned 16-byte chunks starting at the image's first byte
                    f = f.clone()
                keep.append(f)
                ptrs.append(f.data_ptr())
        elif flat:
            sizes = [f.numel() for f in flat]
            offs = np.concatenate([[0], np.cumsum([(n + 15) // 16 * 16 for n in sizes])]).tolist()
            host = torch.empty(offs[-1], dtype=torch.uint8, pin_memory=True)
            for f, o, n in zip(flat, offs, sizes):
                host[o:o + n] = f.reshape(-1).cpu()
            stage = host.to(self.device, non_blocking=True)
            keep.append(stage)
            ptrs = [stage.data_ptr() + o for o in offs[:-1]]
        n_masks = sum(len(ms) for ms in masks)
        out_sam = torch.empty(B, 3, self.sam_size, self.sam_size, dtype=self.out_dtype, device=self.device)
        out_clip = torch.empty(B, 3, self.clip_size, self.clip_size, dtype=self.out_dtype, device=self.device)
        out_mask = torch.empty(n_masks, 1, self.clip_size, self.clip_size, dtype=torch.uint8, device=self.device)
        jobs, resize_list = [], []
        for i, a in enumerate(arrays):
            H, W = int(a.shape[0]), int(a.shape[1])
            j, resize = self._job(ptrs[i], H, W, 3, self.sam_size, self.sam_lut, (0.0, 0.0, 0.0), out_sam[i])
            jobs.append(j)
            resize_list.append(resize)
            jobs.append(self._job(ptrs[i], H, W, 3, self.clip_size, self.clip_lut, self.clip_pad, out_clip[i])[0])
        k = 0
        for i, ms in enumerate(masks):
            H, W = int(arrays[i].shape[0]), int(arrays[i].shape[1])
            for _ in ms:
                jobs.append(self._job(ptrs[B + k], H, W, 1, self.clip_size, None, (0.0, 0.0, 0.0), out_mask[k])[0])
                k += 1
        n_enc = len(enc_flat)
        out_enc = torch.empty(n_enc, 1, self.clip_size, self.clip_size, dtype=self.out_dtype, device=self.device)
        for t, m in enumerate(enc_flat):
            jobs.append(self._job(ptrs[B + n_masks + t], int(m.shape[0]), int(m.shape[1]), 1, self.clip_size,
                                  self.enc_lut, (0.0, 0.0, 0.0), out_enc[t])[0])
        n = len(jobs)
        jobs_host = jobs_dev = None
        if n:
            jobs_host = (PreprocessJob * n)(*jobs)
            blob = torch.frombuffer(bytearray(bytes(jobs_host)), dtype=torch.uint8)
            jobs_dev = blob.pin_memory().to(self.device, non_blocking=True)
        out = {"images": out_sam, "images_clip": out_clip, "resize_list": resize_list}
        if region_masks is not None:
            grids = out_mask[:, :, ::self.patch, ::self.patch]  # cv2.resize(fx=1/14, INTER_NEAREST): every 14th pixel
            out["region_masks_u8"] = out_mask
            out["region_masks"], k = [], 0
            for ms in masks:
                out["region_masks"].append([grids[k + t] for t in range(len(ms))])
                k += len(ms)
        if encoder_masks is not None:  # the collator's `mask_images`: one [m, 1, S, S] tensor per sample that has masks
            out["mask_images"], t = [], 0
            for ms in enc:
                if ms:
                    out["mask_images"].append(out_enc[t:t + len(ms)])
                t += len(ms)
        src_bytes = sum(int(a.shape[0]) * int(a.shape[1]) * 3 * 2 for a in arrays) + sum(m.numel() for ms in masks for m in ms)
        out_bytes = out_sam.numel() * out_sam.element_size() + out_clip.numel() * out_clip.element_size() + out_mask.numel()
        return {"n": n, "jobs_host": jobs_host, "jobs_dev": jobs_dev, "keep": keep, "out": out,
                "algorithmic_bytes": src_bytes + out_bytes}

    def launch(self, plan):
        """ONE kernel launch on the current stream for every job of the plan; returns the output dict."""
        if plan["n"]:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self.lib.mpl_preprocess_images(plan["jobs_host"], ctypes.c_void_p(plan["jobs_dev"].data_ptr()),
                                                      plan["n"], ctypes.c_void_p(stream)), "mpl_preprocess_images")
        return plan["out"]

    def __call__(self, images, region_masks=None, encoder_masks=None):
        return self.launch(self.plan(images, region_masks, encoder_masks))
