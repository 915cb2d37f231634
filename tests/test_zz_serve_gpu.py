"""GPU: the streaming serving loop (medplib_b200/serve.py) against evaluate() on the same small model —
same greedy tokens, same mask — and the ICL exemplar-mask path of the input pipeline kernel.  (First run on a B200 in
round 2, call 1: both green; the round-1 gate is gone.)  The loop's host logic is also covered on CPU by
tests/test_serve_cpu.py."""
import math
import os

import pytest
import torch

from test_model_gpu import SEG, build, inputs

pytestmark = [pytest.mark.gpu]


class Tok:
    pad_token_id, eos_token_id = 0, 2

    def __call__(self, s, add_special_tokens=True):
        raise AssertionError("no stop string in this test")

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(str(i) for i in ids if not (skip_special_tokens and i in (0, 2)))


def test_stream_equals_evaluate(dev):
    from medplib_b200 import serve
    m, _, _ = build(dev)
    ids, clip_img, sam_img = inputs()
    label = torch.zeros(70, 90)
    forced = {3: SEG, 5: 2}
    out_ids, masks = m.evaluate(clip_img.to(dev), sam_img.to(dev), ids.to(dev), [(256, 256)], [label],
                                max_new_tokens=6, forced_tokens=forced)
    recs = list(serve.generate_stream(m, Tok(), ids.to(dev), images_clip=clip_img.to(dev), images_sam=sam_img.to(dev),
                                      resize=(256, 256), original_size=(70, 90), temperature=0.0, max_new_tokens=6,
                                      forced_tokens=forced))
    new = out_ids[0, ids.shape[1]:].tolist()
    assert len(recs) == 6 and recs[-1]["text"] == Tok().decode(new)
    assert all(r["mask"] == [] for r in recs[:-1])
    want = masks[0][0].float()
    thr = math.log(0.1 / 0.9)
    got = torch.zeros(70, 90, dtype=torch.bool)
    for r, c in recs[-1]["mask"]:
        got[r, c] = True
    far = ((want - thr).abs() > 2e-2 * want.abs().max()).cpu()
    assert (recs[-1]["height"], recs[-1]["width"]) == ("70", "90")
    assert torch.equal(got[far], (want.cpu() > thr)[far])


def test_icl_encoder_masks_match_reference_digests(dev):
    """ImagePreprocessor(encoder_masks=...) -> `mask_images` vs the reference's _preprocess_encoder_mask (golden digests).
    Uses the single-channel + value-table kernel path, which the validated tests do not exercise."""
    import hashlib
    import sys
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import inputs as gi
    from oracle import preprocess as op
    from medplib_b200.preprocess import ImagePreprocessor
    gold = torch.load(os.path.join(here, "golden", "preprocess.pt"), weights_only=False)["encoder_masks"]
    idxs = [i for i, s in enumerate(gi.PREPROCESS_SIZES) if s[3] == 336]
    masks = [gi.preprocess_mask(i, *gi.PREPROCESS_SIZES[i][:2]) for i in idxs]
    img = gi.preprocess_image(0, 64, 64)
    out = ImagePreprocessor(dev)([img, img], encoder_masks=[masks[:3], masks[3:]])
    got = torch.cat(out["mask_images"], 0).cpu().numpy()
    assert out["mask_images"][0].shape == (3, 1, 336, 336)
    for t, i in enumerate(idxs):
        assert np.array_equal(got[t], op.encoder_mask(masks[t])), f"case {i}"
        assert hashlib.sha256(np.ascontiguousarray(got[t]).tobytes()).hexdigest() == gold[i]["sha256"]
