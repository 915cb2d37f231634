"""CPU: the input-pipeline oracle (oracle/preprocess.py) against the reference's own pipeline — the committed
digests in tests/golden/preprocess.pt (made by tests/golden/make_golden_preprocess.py from the reference's
ResizeLongestSide / LazySupervisedDataset.preprocess / CLIPImageProcessor / cv2) and, when PIL is importable, PIL live."""
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

GOLD = torch.load(os.path.join(HERE, "golden", "preprocess.pt"), weights_only=False)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def oracle_outputs(i, h, w, ls, lc):
    img = gi.preprocess_image(i, h, w)
    sam_u8 = op.resize_longest_side(img, ls)
    image_sam, resize = op.image_sam(img, ls)
    rc = op.resize_longest_side(img, lc)
    clip_u8 = op.centre_pad(np.ascontiguousarray(rc.transpose(2, 0, 1)), lc, op.CLIP_PAD_U8)
    m = gi.preprocess_mask(i, h, w)
    region_u8 = op.centre_pad(op.resize_longest_side(m, lc)[None], lc, np.zeros(1, np.uint8))[0]
    return {"sam_u8": sam_u8, "image_sam": image_sam, "clip_u8": clip_u8, "image_clip": op.image_clip(img, lc),
            "region_u8": region_u8, "region_grid": op.region_mask(m, lc)}, resize


@pytest.mark.parametrize("i", range(len(gi.PREPROCESS_SIZES)))
def test_oracle_matches_reference_pipeline_bit_for_bit(i):
    h, w, ls, lc = gi.PREPROCESS_SIZES[i]
    case = GOLD["cases"][i]
    assert case["hw"] == (h, w) and case["targets"] == (ls, lc)
    outs, resize = oracle_outputs(i, h, w, ls, lc)
    assert tuple(resize) == case["resize"]
    for name, arr in outs.items():
        if "tensors" in case:  # small case kept whole: report where, not just that
            ref = case["tensors"][name].numpy()
            assert arr.shape == ref.shape and arr.dtype == ref.dtype, name
            assert np.array_equal(arr, ref), f"{name}: {np.argwhere(arr != ref)[:4]}"
        assert digest(arr) == case["sha256"][name], f"case {i} {name}: differs from the reference's output"
    assert np.array_equal(outs["region_grid"], case["region_grid"].numpy())


def test_coefficient_tables_are_normalised_and_monotone():
    for n_in, n_out in [(512, 256), (97, 191), (1777, 256), (336, 336), (40, 269), (2, 336), (4000, 7)]:
        bounds, coeffs = op.pil_coeffs(n_in, n_out)
        assert (np.abs(coeffs.sum(1) - (1 << op.PRECISION_BITS)) <= coeffs.shape[1]).all()
        assert (np.diff(bounds[:, 0]) >= 0).all() and (bounds[:, 1] >= 1).all()
        assert (bounds[:, 0] + bounds[:, 1] <= n_in).all()
        # the band bound the CUDA kernel sizes its shared memory with (medplib_b200/preprocess.py:band_rows_bound)
        from medplib_b200.preprocess import band_rows_bound
        for R in (1, 2, 4, 8):
            for y0 in range(0, n_out, R):
                y1 = min(y0 + R, n_out)
                rows = bounds[y1 - 1, 0] + bounds[y1 - 1, 1] - bounds[y0, 0]
                assert rows <= band_rows_bound(n_in, n_out, R), (n_in, n_out, R, y0)


def test_constant_image_and_identity():
    img = np.full((123, 77, 3), 201, np.uint8)
    assert (op.resize_longest_side(img, 256) == 201).all()  # weights sum to one within rounding
    img = gi.preprocess_image(0, 200, 336)
    assert np.array_equal(op.resize_longest_side(img, 336), img)  # both passes skipped
    sam, resize = op.image_sam(img, 256)
    assert resize == (152, 256) and sam.shape == (3, 256, 256)
    assert (sam[:, :52] == 0).all() and (sam[:, 52 + 152:] == 0).all()  # (256 - 152) // 2 rows of zero padding on top


def test_live_pil_when_available():
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    for (h, w, nh, nw) in [(301, 203, 256, 173), (33, 900, 12, 336), (700, 700, 336, 336), (5, 5, 256, 256)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.array(PIL.fromarray(img).resize((nw, nh), PIL.BILINEAR))
        assert np.array_equal(op.pil_resize_bilinear(img, nh, nw), ref)


def test_product_host_tables_equal_the_oracle():
    """medplib_b200/preprocess.py builds the kernel's tables on the host (vectorised); they must equal the oracle's
    sequential restatement of PIL for every size pair, and the value tables must equal the oracle's."""
    from medplib_b200 import preprocess as pp
    rng = np.random.default_rng(1)
    pairs = [(512, 256), (512, 336), (97, 191), (1777, 256), (336, 336), (40, 269), (2, 336), (4000, 7), (1, 5), (5, 1)]
    pairs += [(int(a), int(b)) for a, b in rng.integers(1, 3000, (60, 2))]
    for n_in, n_out in pairs:
        b0, c0 = op.pil_coeffs(n_in, n_out)
        b1, c1 = pp.pil_coeffs(n_in, n_out)
        assert np.array_equal(b0, b1) and np.array_equal(c0, c1), (n_in, n_out)
    assert np.array_equal(pp.sam_level_table().numpy(), op.sam_lut())
    assert np.array_equal(pp.clip_level_table().numpy(), op.clip_lut())
    assert pp.clip_pad_levels() == op.CLIP_PAD_U8.tolist() == [122, 116, 104]
    for hw in [(512, 512), (300, 451), (1, 9), (1300, 1777)]:
        for L in (256, 336):
            assert pp.get_preprocess_shape(*hw, L) == op.get_preprocess_shape(*hw, L)


def test_library_band_bound_matches_host_mirror():
    from medplib_b200 import _lib
    from medplib_b200.preprocess import band_rows_bound
    lib = _lib.load()
    for n_in, n_out in [(512, 256), (97, 191), (1777, 256), (336, 336), (40000, 256), (3, 336)]:
        for R in (1, 2, 4, 8):
            assert lib.mpl_preprocess_band_rows(n_in, n_out, R) == band_rows_bound(n_in, n_out, R)
    assert lib.mpl_preprocess_band_rows(0, 4, 1) == -1


def test_preprocessor_fails_loudly_without_gpu():
    from medplib_b200 import _lib
    from medplib_b200.preprocess import ImagePreprocessor
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MplError):
        ImagePreprocessor()


def test_oracle_equals_pil_on_random_shapes():
    """Property check over the size space (up- and down-scales, 1-pixel axes, non-integer ratios, both channel counts)."""
    PIL = pytest.importorskip("PIL.Image")
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import strategies as st

    @hyp.settings(max_examples=60, deadline=None, derandomize=True)
    @hyp.given(st.integers(1, 90), st.integers(1, 90), st.integers(1, 120), st.integers(1, 120), st.booleans(),
               st.integers(0, 2 ** 31 - 1))
    def check(h, w, nh, nw, rgb, seed):
        rng = np.random.default_rng(seed)
        img = rng.integers(0, 256, (h, w, 3) if rgb else (h, w), dtype=np.uint8)
        ref = np.array(PIL.fromarray(img).resize((nw, nh), PIL.BILINEAR))
        assert np.array_equal(op.pil_resize_bilinear(img, nh, nw), ref), (h, w, nh, nw, rgb)

    check()


def test_icl_encoder_masks_match_the_reference():
    """ICLLazySupervisedDataset._preprocess_encoder_mask digests (golden) vs the oracle."""
    assert len(GOLD["encoder_masks"]) == len(gi.PREPROCESS_SIZES)
    for i, (h, w, ls, lc) in enumerate(gi.PREPROCESS_SIZES):
        em = op.encoder_mask(gi.preprocess_mask(i, h, w), lc)
        assert em.shape == (1, lc, lc) and em.dtype == np.float32
        assert int(em.sum()) == GOLD["encoder_masks"][i]["ones"]
        assert digest(em) == GOLD["encoder_masks"][i]["sha256"], f"case {i}"


def _emulate_job(j):
    """What csrc/preprocess.cu computes for one mpl_preprocess_job, read straight from the struct's (host) pointers."""
    import ctypes

    def arr(ptr, n, ct, dt):
        return np.frombuffer((ct * n).from_address(ptr), dtype=dt).copy()
    H, W, C, nh, nw, L = j.H, j.W, j.C, j.new_h, j.new_w, j.out_size
    assert j.src % 16 == 0 and j.ks_x % 4 == 0 and j.src_stride == W * C
    src = arr(j.src, H * W * C, ctypes.c_ubyte, np.uint8).reshape(H, W, C).astype(np.int64)
    bx = arr(j.bound_x, nw * 2, ctypes.c_int, np.int32).reshape(nw, 2)
    cx = arr(j.coef_x, j.ks_x * nw, ctypes.c_int, np.int32).reshape(j.ks_x, nw).astype(np.int64)
    by = arr(j.bound_y, nh * 2, ctypes.c_int, np.int32).reshape(nh, 2)
    cy = arr(j.coef_y, nh * j.ks_y, ctypes.c_int, np.int32).reshape(nh, j.ks_y).astype(np.int64)
    tmp = np.zeros((H, nw, C), np.int64)
    for xx in range(nw):
        x0, n = bx[xx]
        acc = (1 << 21) + (src[:, x0:x0 + n, :] * cx[:n, xx][None, :, None]).sum(1)
        tmp[:, xx] = np.clip(acc >> 22, 0, 255)
    lv = np.zeros((nh, nw, C), np.int64)
    for yy in range(nh):
        y0, n = by[yy]
        acc = (1 << 21) + (tmp[y0:y0 + n] * cy[yy, :n][:, None, None]).sum(0)
        lv[yy] = np.clip(acc >> 22, 0, 255)
    if j.lut:
        lut = arr(j.lut, C * 256, ctypes.c_float, np.float32).reshape(C, 256)
        val = np.stack([lut[c][lv[..., c]] for c in range(C)])
    else:
        val = lv.transpose(2, 0, 1).astype(np.float32)
    out = np.empty((C, L, L), np.float32)
    out[:] = np.array(list(j.pad_value)[:C], np.float32).reshape(C, 1, 1)
    out[:, j.pad_top:j.pad_top + nh, j.pad_left:j.pad_left + nw] = val
    return out


def test_host_side_job_structs_describe_the_oracle_result(monkeypatch):
    """ImagePreprocessor.plan() on CPU tensors (pinning stubbed out): the job array it hands to mpl_preprocess_images —
    source offsets in the ragged staging buffer, sizes, pads, PIL tables in the kernel's layouts, value tables — evaluated
    by a numpy reading of the kernel's contract, equals the oracle for images, region masks and ICL encoder masks."""
    from medplib_b200 import preprocess as pp
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "pin_memory"}))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    pre = object.__new__(pp.ImagePreprocessor)
    pre.device = torch.device("cpu")
    pre.sam_size, pre.clip_size, pre.patch, pre.out_dtype = 64, 84, 14, torch.float32
    pre.sam_lut, pre.clip_lut = pp.sam_level_table().contiguous(), pp.clip_level_table().contiguous()
    pre.clip_pad = [float(pre.clip_lut[c, lvl]) for c, lvl in enumerate(pp.clip_pad_levels())]
    pre.enc_lut = (torch.arange(256) > 0).float().view(1, 256).contiguous()
    pre._tables = {}
    imgs = [gi.preprocess_image(i, h, w) for i, (h, w) in enumerate([(50, 40), (211, 97), (64, 84)])]
    rms = [[gi.preprocess_mask(i, *im.shape[:2])] for i, im in enumerate(imgs)]
    ems = [[gi.preprocess_mask(7, 33, 90)], [], [gi.preprocess_mask(8, 120, 45), gi.preprocess_mask(9, 84, 84)]]
    plan = pre.plan(imgs, region_masks=rms, encoder_masks=ems)
    jobs = list(plan["jobs_host"])
    assert plan["n"] == len(jobs) == 2 * 3 + 3 + 3
    for b, im in enumerate(imgs):
        assert np.array_equal(_emulate_job(jobs[2 * b]), op.image_sam(im, 64)[0])
        assert np.array_equal(_emulate_job(jobs[2 * b + 1]), op.image_clip(im, 84))
        want = op.centre_pad(op.resize_longest_side(rms[b][0], 84)[None], 84, np.zeros(1, np.uint8))
        assert np.array_equal(_emulate_job(jobs[6 + b]), want.astype(np.float32))
    flat = [m for ms in ems for m in ms]
    for t, m in enumerate(flat):
        assert np.array_equal(_emulate_job(jobs[9 + t]), op.encoder_mask(m, 84))
    assert [tuple(x.shape) for x in plan["out"]["mask_images"]] == [(1, 1, 84, 84), (2, 1, 84, 84)]
    assert plan["out"]["resize_list"] == [op.get_preprocess_shape(*im.shape[:2], 64) for im in imgs]


# ------------------------------------------------------------------------------------------------------------------
# The kernel source itself, executed on the CPU (tests/dev/cuda_emu.h: one std::thread per CUDA thread, std::barrier for
# __syncthreads, checked memcpy for cp.async): index arithmetic, shared-memory layout, staging, clamping and both loop
# structures against the oracle — without a GPU.  The B200 run (tests/test_preprocess_gpu.py) remains the parity test.
def _cpu_preprocessor(monkeypatch, sam, clip, out_dtype=torch.float32):
    from medplib_b200 import preprocess as pp
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "pin_memory"}))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    pre = object.__new__(pp.ImagePreprocessor)
    pre.device = torch.device("cpu")
    pre.sam_size, pre.clip_size, pre.patch, pre.out_dtype = sam, clip, 14, out_dtype
    pre.sam_lut, pre.clip_lut = pp.sam_level_table().contiguous(), pp.clip_level_table().contiguous()
    pre.clip_pad = [float(pre.clip_lut[c, lvl]) for c, lvl in enumerate(pp.clip_pad_levels())]
    pre.enc_lut = (torch.arange(256) > 0).float().view(1, 256).contiguous()
    pre._tables = {}
    return pre


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    import ctypes
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(HERE)
    out = str(tmp_path_factory.mktemp("emu") / "preprocess_emu.so")
    cmd = ["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", "-DMPL_CPU_EMULATION", "-I",
           os.path.join(root, "tests", "dev"), "-x", "c++", os.path.join(root, "tests", "dev", "preprocess_emu.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = ctypes.CDLL(out)
    lib.mpl_emu_preprocess_images.restype = ctypes.c_int
    return lib


@pytest.mark.parametrize("variant", [0, 1], ids=["default_loops", "v4_candidate"])
def test_kernel_source_on_cpu_matches_the_oracle(emu_lib, monkeypatch, variant):
    pre = _cpu_preprocessor(monkeypatch, 64, 84)
    sizes = [(50, 40), (211, 97), (64, 84), (31, 200), (300, 9), (1100, 130)]
    imgs = [gi.preprocess_image(i, h, w) for i, (h, w) in enumerate(sizes)]
    rms = [[gi.preprocess_mask(i, *im.shape[:2])] for i, im in enumerate(imgs)]
    ems = [[gi.preprocess_mask(7, 33, 90)], [], [gi.preprocess_mask(8, 120, 45)], [], [], []]
    plan = pre.plan(imgs, region_masks=rms, encoder_masks=ems)
    for t in plan["out"].values():
        if torch.is_tensor(t):
            t.fill_(-7)
    assert emu_lib.mpl_emu_preprocess_images(plan["jobs_host"], plan["n"], variant) == 0, \
        "the kernel read outside the ranges the C ABI allows (or the jobs were rejected)"
    out = plan["out"]
    for b, im in enumerate(imgs):
        assert np.array_equal(out["images"][b].numpy(), op.image_sam(im, 64)[0]), f"images[{b}] {im.shape}"
        assert np.array_equal(out["images_clip"][b].numpy(), op.image_clip(im, 84)), f"images_clip[{b}] {im.shape}"
        want = op.centre_pad(op.resize_longest_side(rms[b][0], 84)[None], 84, np.zeros(1, np.uint8))
        assert np.array_equal(out["region_masks_u8"][b].numpy(), want), f"region mask {b}"
    got = torch.cat(out["mask_images"], 0).numpy()
    for t, m in enumerate([m for ms in ems for m in ms]):
        assert np.array_equal(got[t], op.encoder_mask(m, 84)), f"encoder mask {t}"


def test_kernel_source_on_cpu_full_size_and_bf16(emu_lib, monkeypatch):
    """The real targets (256 / 336), a 4x downscale with the 4-row stage, bf16 output: both loop structures."""
    pre = _cpu_preprocessor(monkeypatch, 256, 336, torch.bfloat16)
    im = gi.preprocess_image(3, 700, 1024)
    want_s = torch.from_numpy(op.image_sam(im)[0]).to(torch.bfloat16)
    want_c = torch.from_numpy(op.image_clip(im)).to(torch.bfloat16)
    for variant in (0, 1):
        plan = pre.plan([im])
        assert emu_lib.mpl_emu_preprocess_images(plan["jobs_host"], plan["n"], variant) == 0
        assert torch.equal(plan["out"]["images"][0], want_s), variant
        assert torch.equal(plan["out"]["images_clip"][0], want_c), variant
