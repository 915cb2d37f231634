"""Oracle: image input pipeline (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).  SURVEY §8 f-1.

numpy restatement of what datasets/LazySupervisedDataset.py:535-556 does to one decoded RGB image between
``cv2.cvtColor`` and the collator:

  * ``ResizeLongestSide.apply_image`` (model/segment_anything_med2d/utils/transforms.py:26-32, ``get_preprocess_shape``
    :95-103) = torchvision ``resize(to_pil_image(img), (newh, neww))`` = PIL ``Image.resize(BILINEAR)``.  The arithmetic
    lives in the un-vendored Pillow (libImaging/Resample.c, any release since 3.4; 12.2.0 installed here): an
    antialiased two-pass (horizontal, then vertical) convolution on 8-bit channels with 22-bit fixed-point
    coefficients and a u8 rounding between the passes.  ``pil_coeffs`` / ``pil_resize_bilinear`` restate it.
  * ``preprocess`` + ``pad_tensor_channelwise`` (LazySupervisedDataset.py:446-502): SAM branch = (x - mean) / std in fp32
    then centre zero-pad to 256; CLIP branch = centre pad of the u8 image with int(clip_mean * 255) to 336, then
    ``CLIPImageProcessor.preprocess`` (transformers 4.31: resize / centre-crop are the identity on a 336x336 input,
    ``rescale`` = u8 * (1/255) in float64 cast to fp32, ``normalize`` = (x - mean_f32) / std_f32).
  * region masks (LazySupervisedDataset.py:516-519): the same resize on the single-channel u8 mask, zero pad to 336,
    ``cv2.resize(fx=fy=1/14, INTER_NEAREST)`` = every 14th pixel -> 24x24.

Pinned: tests/golden/preprocess.pt holds outputs of the reference's own ResizeLongestSide + PIL / torchvision / cv2 /
CLIPImageProcessor run in the authoring container (tests/golden/make_golden_preprocess.py); tests/test_preprocess_cpu.py
checks this file against them bit for bit (u8) and to the last fp32 bit (normalised tensors).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c

SAM_MEAN = np.array([123.675, 116.28, 103.53], np.float32)  # LazySupervisedDataset.py:394-395
SAM_STD = np.array([58.395, 57.12, 57.375], np.float32)
CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], np.float32)  # OPENAI_CLIP_MEAN / STD
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], np.float32)
# LazySupervisedDataset.py:398: (mean * 255).clamp(0, 255).to(torch.int) on fp32 -> truncation
CLIP_PAD_U8 = (CLIP_MEAN * np.float32(255)).clip(0, 255).astype(np.int32).astype(np.uint8)


def get_preprocess_shape(oldh, oldw, long_side):
    """transforms.py:95-103."""
    scale = long_side * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


def pil_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (support 1.0) over the whole axis.
    Returns (bounds int32 [out, 2] = (first tap, tap count), coeffs int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        x = np.arange(xmax, dtype=np.float64)
        w = np.maximum(1.0 - np.abs((x + xmin - center + 0.5) * ss), 0.0)
        ww = 0.0
        for v in w:  # same summation order as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    coeffs = np.where(kk < 0, -0.5 + kk * (1 << PRECISION_BITS), 0.5 + kk * (1 << PRECISION_BITS)).astype(np.int32)
    return bounds, coeffs


def _resample_axis0(img, out_size):
    """One 8bpc pass along axis 0 of a u8 array [n, ...]."""
    n = img.shape[0]
    if out_size == n:
        return img  # ImagingResample skips a pass whose size does not change
    bounds, coeffs = pil_coeffs(n, out_size)
    ksize = coeffs.shape[1]
    idx = np.minimum(bounds[:, :1] + np.arange(ksize)[None, :], n - 1)  # taps past the count carry zero weight
    acc = np.full((out_size,) + img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
    src = img.astype(np.int64)
    for t in range(ksize):
        acc += src[idx[:, t]] * coeffs[:, t].reshape((-1,) + (1,) * (img.ndim - 1))
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bilinear(img, new_h, new_w):
    """PIL Image.resize((new_w, new_h), BILINEAR) of a u8 array [H, W] or [H, W, C]: horizontal pass, then vertical."""
    assert img.dtype == np.uint8
    tmp = np.swapaxes(_resample_axis0(np.swapaxes(img, 0, 1), new_w), 0, 1)
    return np.ascontiguousarray(_resample_axis0(tmp, new_h))


def resize_longest_side(img, long_side):
    """ResizeLongestSide(long_side).apply_image."""
    new_h, new_w = get_preprocess_shape(img.shape[0], img.shape[1], long_side)
    return pil_resize_bilinear(img, new_h, new_w)


def centre_pad(x, size, values):
    """pad_tensor_channelwise on a [C, h, w] array: total pad split floor / rest, per-channel fill."""
    c, h, w = x.shape
    top, left = (size - h) // 2, (size - w) // 2
    out = np.empty((c, size, size), x.dtype)
    out[:] = np.asarray(values, x.dtype).reshape(c, 1, 1)
    out[:, top:top + h, left:left + w] = x
    return out


def sam_lut():
    """fp32 value of every u8 level per channel on the SAM branch: (x - mean) / std in fp32."""
    lv = np.arange(256, dtype=np.float32)[None, :]
    return ((lv - SAM_MEAN[:, None]) / SAM_STD[:, None]).astype(np.float32)


def clip_lut():
    """fp32 value of every u8 level per channel through CLIPImageProcessor (transformers 4.31 rescale + normalize)."""
    lv = (np.arange(256, dtype=np.uint8)[None, :] * 0.00392156862745098).astype(np.float32)
    return ((lv - CLIP_MEAN[:, None]) / CLIP_STD[:, None]).astype(np.float32)


def image_sam(image_rgb, sam_size=256):
    """-> (fp32 [3, S, S], resize (h, w)) : LazySupervisedDataset.py:539-541."""
    r = resize_longest_side(image_rgb, sam_size)
    lut = sam_lut()
    x = np.stack([lut[c][r[..., c]] for c in range(3)])
    return centre_pad(x, sam_size, np.zeros(3, np.float32)), r.shape[:2]


def image_clip(image_rgb, clip_size=336):
    """-> fp32 [3, 336, 336] : LazySupervisedDataset.py:546-553 (image_aspect_ratio == 'pad')."""
    r = resize_longest_side(image_rgb, clip_size)
    u8 = centre_pad(np.ascontiguousarray(r.transpose(2, 0, 1)), clip_size, CLIP_PAD_U8)
    lut = clip_lut()
    return np.stack([lut[c][u8[c]] for c in range(3)])


def region_mask(mask_u8, clip_size=336, patch=14):
    """-> u8 [24, 24] : LazySupervisedDataset.py:516-519 up to (not including) generate_mask_with_sub_component."""
    r = resize_longest_side(mask_u8, clip_size)
    padded = centre_pad(r[None], clip_size, np.zeros(1, np.uint8))[0]
    return np.ascontiguousarray(padded[::patch, ::patch])


def encoder_mask(mask_u8, clip_size=336):
    """-> fp32 {0,1} [1, S, S] : ICLLazySupervisedDataset._preprocess_encoder_mask (datasets/ICLLazySupervisedDataset.py
    :77-85): the {0,1} mask times 255, the same resize and zero pad, then ``> 0``."""
    r = resize_longest_side((mask_u8 != 0).astype(np.uint8) * 255, clip_size)
    return (centre_pad(r[None], clip_size, np.zeros(1, np.uint8)) > 0).astype(np.float32)
