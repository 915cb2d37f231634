"""GPU: the one-launch image input pipeline (medplib_b200/preprocess.py -> mpl_preprocess_images) against
the oracle and the reference's own outputs (tests/golden/preprocess.pt) — bit-exact: u8 resampling is integer work and
every fp32 output value is a table entry."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import inputs as gi  # noqa: E402

from oracle import preprocess as op  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(HERE, "golden", "preprocess.pt"), weights_only=False)


def digest(t):
    return hashlib.sha256(np.ascontiguousarray(t.cpu().numpy()).tobytes()).hexdigest()


def test_ragged_batches_match_reference_digests_and_oracle(dev):
    from medplib_b200.preprocess import ImagePreprocessor
    groups = {}
    for i, (h, w, ls, lc) in enumerate(gi.PREPROCESS_SIZES):
        groups.setdefault((ls, lc), []).append(i)
    for (ls, lc), idxs in groups.items():
        pre = ImagePreprocessor(dev, sam_size=ls, clip_size=lc)
        imgs = [gi.preprocess_image(i, *gi.PREPROCESS_SIZES[i][:2]) for i in idxs]
        masks = [[gi.preprocess_mask(i, *gi.PREPROCESS_SIZES[i][:2])] for i in idxs]
        out = pre(imgs, region_masks=masks)
        assert out["images"].shape == (len(idxs), 3, ls, ls) and out["images"].dtype == torch.float32
        assert out["images_clip"].shape == (len(idxs), 3, lc, lc)
        for b, i in enumerate(idxs):
            case = GOLD["cases"][i]
            assert tuple(out["resize_list"][b]) == case["resize"]
            sam, _ = op.image_sam(imgs[b], ls)
            assert np.array_equal(out["images"][b].cpu().numpy(), sam), f"case {i}: images differs from the oracle"
            assert np.array_equal(out["images_clip"][b].cpu().numpy(), op.image_clip(imgs[b], lc)), f"case {i}: images_clip"
            assert digest(out["images"][b]) == case["sha256"]["image_sam"]
            assert digest(out["images_clip"][b]) == case["sha256"]["image_clip"]
            assert digest(out["region_masks_u8"][b, 0]) == case["sha256"]["region_u8"]
            assert torch.equal(out["region_masks"][b][0][0].cpu(), case["region_grid"])


def test_bf16_output_device_inputs_and_large_downscale(dev):
    from medplib_b200.preprocess import ImagePreprocessor
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, s, dtype=np.uint8) for s in [(3000, 4000, 3), (8, 700, 3), (640, 5, 3), (2048, 2048, 3)]]
    f32 = ImagePreprocessor(dev)(imgs)
    bf = ImagePreprocessor(dev, out_dtype=torch.bfloat16)([torch.from_numpy(a).to(dev) for a in imgs])  # device-resident
    for b, a in enumerate(imgs):
        sam, resize = op.image_sam(a)
        assert tuple(f32["resize_list"][b]) == tuple(resize)
        assert np.array_equal(f32["images"][b].cpu().numpy(), sam), a.shape
        assert np.array_equal(f32["images_clip"][b].cpu().numpy(), op.image_clip(a)), a.shape
    assert torch.equal(bf["images"], f32["images"].to(torch.bfloat16))
    assert torch.equal(bf["images_clip"], f32["images_clip"].to(torch.bfloat16))
    with pytest.raises(ValueError):
        ImagePreprocessor(dev)([np.zeros((1, 700, 3), np.uint8)])


def test_feeds_the_model_contract_and_counts_one_launch(dev):
    """Output keys / shapes / dtypes are the collator's (a-0) and the whole batch is ONE kernel launch."""
    from medplib_b200 import _lib
    from medplib_b200.preprocess import ImagePreprocessor
    pre = ImagePreprocessor(dev)
    imgs = [gi.preprocess_image(i, 512 + 64 * i, 700 - 50 * i) for i in range(8)]
    pre(imgs)  # tables cached
    torch.cuda.synchronize()
    n0 = _lib.load().mpl_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pre(imgs)
    e1.record()
    torch.cuda.synchronize()
    assert _lib.load().mpl_launch_count() - n0 == 1
    assert out["images"].shape == (8, 3, 256, 256) and out["images_clip"].shape == (8, 3, 336, 336)
    assert out["images"].is_contiguous() and out["images"].device.type == "cuda"
    os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(HERE), "gpurun_out", "preprocess_timing.txt"), "a") as f:
        f.write(f"batch 8 ragged ~512..960 px, host u8 in, fp32 out, incl. H2D + host job build: {e0.elapsed_time(e1):.3f} ms\n")
