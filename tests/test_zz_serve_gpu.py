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


def test_continuous_batching_equals_single_requests(dev):
    """serve.ContinuousBatcher: three requests of different lengths (one with an image + <SEG> -> mask tail), the third
    submitted while the first two are decoding, against the SAME requests served one at a time by generate_stream.
    Every request follows forced tokens (random weights), so both runs walk the same trajectory; compared are the text
    records and the mask. The batched run uses ONE persistent decode kernel launch per token step for all active slots
    with per-sequence RoPE positions (mpl_llama_io.rope_pos) behind a shared cache column."""
    from medplib_b200 import serve
    m, _, _ = build(dev)
    ids, clip_img, sam_img = inputs()
    g = torch.Generator().manual_seed(3)
    text_a = torch.randint(3, 290, (1, 9), generator=g)
    text_b = torch.randint(3, 290, (1, 17), generator=g)
    reqs = [
        dict(input_ids=ids.to(dev), images_clip=clip_img.to(dev), images_sam=sam_img.to(dev), resize=(256, 256),
             original_size=(70, 90), max_new_tokens=7, forced_tokens={0: 11, 1: 12, 2: 13, 3: SEG, 4: 14, 5: 15, 6: 2}),
        dict(input_ids=text_a.to(dev), max_new_tokens=5, forced_tokens={0: 21, 1: 22, 2: 23, 3: 24, 4: 2}),
        dict(input_ids=text_b.to(dev), max_new_tokens=9, forced_tokens={i: 31 + i for i in range(8)} | {8: 2}),
    ]
    single = []
    for r in reqs:
        recs = list(serve.generate_stream(m, Tok(), r["input_ids"], images_clip=r.get("images_clip"),
                                          images_sam=r.get("images_sam"), resize=r.get("resize"),
                                          original_size=r.get("original_size"), temperature=0.0,
                                          max_new_tokens=r["max_new_tokens"], forced_tokens=r["forced_tokens"]))
        single.append(recs[-1])
    cb = serve.ContinuousBatcher(m, Tok(), max_batch=4, max_len=256, temperature=0.0)
    cb.submit(request_id=0, **reqs[0])
    cb.submit(request_id=1, **reqs[1])
    final = {}
    for _ in range(2):
        for rid, rec, done in cb.step():
            if done:
                final[rid] = rec
    cb.submit(request_id=2, **reqs[2])  # joins while 0 and 1 are mid-flight; its 17-token prompt sits behind the column
    final.update(cb.run())
    assert set(final) == {0, 1, 2}
    for i in range(3):
        assert final[i]["text"] == single[i]["text"], i
    assert (final[0]["height"], final[0]["width"]) == ("70", "90")
    a, b = torch.zeros(70, 90, dtype=torch.bool), torch.zeros(70, 90, dtype=torch.bool)
    for r, c in final[0]["mask"]:
        a[r, c] = True
    for r, c in single[0]["mask"]:
        b[r, c] = True
    # same hidden row up to the summation order of the split-K decode attention: pixels may differ only at the threshold
    assert (a != b).float().mean().item() < 5e-3, "mask of the batched request differs from the single-request mask"
    assert final[1]["mask"] == [] and final[2]["mask"] == []


def test_per_sequence_rope_positions_match_the_plain_decode_step(dev):
    """mpl_llama_io.rope_pos with positions equal to the cache column is the ordinary decode step (bit-identical), and
    a sequence whose keys sit at shifted columns behind a key mask reproduces its un-shifted logits."""
    from medplib_b200 import ops
    m, _, _ = build(dev)
    eng = m._llama()
    g = torch.Generator().manual_seed(5)
    D = m.config.hidden_size
    n, shift = 12, 7
    x = (torch.randn(1, n, D, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    tok = (torch.randn(1, 1, D, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    c0 = eng.new_cache(1, 64)
    eng.forward(x.clone(), c0)
    ref = eng.forward(tok.clone(), c0)["last_hidden_state"].float()
    # same step with an explicit position = column
    c1 = eng.new_cache(1, 64)
    eng.forward(x.clone(), c1)
    got = eng.forward(tok.clone(), c1, rope_pos=torch.tensor([n], dtype=torch.int32, device=dev),
                      kv_mask=torch.ones(1, 64, dtype=torch.uint8, device=dev))["last_hidden_state"].float()
    assert torch.equal(got, ref)
    # keys moved to columns [shift, shift + n) of a wider batch row, masked elsewhere; rope position stays n
    c2 = eng.new_cache(2, 64)
    c2.k[:, 1, :, shift:shift + n] = c1.k[:, 0, :, :n]
    c2.v[:, 1, :, shift:shift + n] = c1.v[:, 0, :, :n]
    mask = torch.zeros(2, 64, dtype=torch.uint8, device=dev)
    mask[1, shift:shift + n] = 1
    mask[:, shift + n] = 1
    c2.len = shift + n
    x2 = torch.cat([tok, tok], 0).clone()
    out = eng.forward(x2, c2, kv_mask=mask, rope_pos=torch.tensor([0, n], dtype=torch.int32, device=dev))
    got2 = out["last_hidden_state"][1].float()
    err = (got2 - ref[0]).abs().max().item() / ref.abs().max().item()
    assert err < 2e-2, err
