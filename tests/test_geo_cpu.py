"""CPU: pins oracle/geo.py (GeoRegionSampler) against the REFERENCE's own module through
tests/golden/geo.pt (made by tests/golden/make_golden_geo.py from /root/reference/model/rp_sampler/GeoSampler.py).

What is held equal, per case: the sampled points (same RNG draws), the FPS indices of every stage (bit-exact: the
first-maximum rule is the reference's), the kNN DISTANCE multisets (the index choice among equal distances is the
implementation's in the reference, see oracle/geo.py), and the final output with the reference's kNN choice injected.
"""
import os

import pytest
import torch

from oracle import geo

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = torch.load(os.path.join(HERE, "golden", "geo.pt"), weights_only=False)
DT = {"torch.float32": torch.float32, "torch.bfloat16": torch.bfloat16}


def unpack(case):
    dt = DT[case["dtype"]]
    d, out_dim, n_init, subs, neighs = case["cfg"]
    sd = {k: v.to(dt) for k, v in case["sd"].items()}
    fmaps = [f.to(dt) for f in case["fmaps"]]
    masks = [[m.long() for m in per] for per in case["masks"]]
    n_stage = len(subs)
    draws = [t for _, t in case["draws"]]
    sample_draws, fps_start = draws[:len(draws) - n_stage], draws[len(draws) - n_stage:]
    return dt, (n_init, subs, neighs), sd, fmaps, masks, sample_draws, fps_start


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_geo_sampler_matches_reference(ci):
    case = CASES[ci]
    dt, (n_init, subs, neighs), sd, fmaps, masks, sample_draws, fps_start = unpack(case)
    rec = {}
    out = geo.geo_region_sampler(sd, "", fmaps, masks, dt, dt, n_init, subs, neighs, case["pooler"],
                                 draws=list(sample_draws), fps_start=fps_start, knn_override=case["knn"], record=rec)
    for s in range(len(subs)):
        assert torch.equal(rec["fps"][s], case["fps"][s]), f"FPS indices of stage {s}"
    tol = 2e-5 if dt == torch.float32 else 2e-2
    for o, r in zip(out, case["out"]):
        assert (o is None) == (r is None)
        if o is not None:
            assert o.shape == r.shape and o.dtype == r.dtype
            scale = float(r.float().abs().max())
            assert float((o.float() - r.float()).abs().max()) <= tol * scale


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_same_seed_draws_the_reference_points(ci):
    """The restatement makes the reference's RNG calls in the reference's order: seeded alike, it needs no injection."""
    case = CASES[ci]
    dt, (n_init, subs, neighs), sd, fmaps, masks, sample_draws, fps_start = unpack(case)
    rec_a, rec_b = {}, {}
    geo.geo_region_sampler(sd, "", fmaps, masks, dt, dt, n_init, subs, neighs, case["pooler"],
                           draws=list(sample_draws), fps_start=fps_start, record=rec_a)
    torch.manual_seed(1234)
    geo.geo_region_sampler(sd, "", fmaps, masks, dt, dt, n_init, subs, neighs, case["pooler"], record=rec_b)
    torch.manual_seed(1234)
    pts = torch.cat([geo.sample_points(m, n_init) for m in masks if len(m)], 0)
    assert torch.equal(pts.to(dt), rec_b["points"])
    assert rec_a["points"].shape == rec_b["points"].shape


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_knn_choice_is_a_valid_topk_of_the_reference(ci):
    """Stable (distance, index) kNN: the same distance multiset as the reference's torch.topk at every anchor, and the
    same index set wherever the k-th and (k+1)-th distances differ."""
    case = CASES[ci]
    dt, (n_init, subs, neighs), sd, fmaps, masks, sample_draws, fps_start = unpack(case)
    rec = {}
    geo.geo_region_sampler(sd, "", fmaps, masks, dt, dt, n_init, subs, neighs, case["pooler"],
                           draws=list(sample_draws), fps_start=fps_start, record=rec)
    xy = rec["points"]
    n_equal = n_total = 0
    for s, k in enumerate(neighs):
        fi = rec["fps"][s]
        assert torch.equal(fi, case["fps"][s])
        new_xy = geo.index_points(xy, fi)
        d = geo.square_distance(new_xy, xy).float()
        mine, ref = rec["knn"][s], case["knn"][s]
        dm, dr = torch.gather(d, -1, mine).sort(-1)[0], torch.gather(d, -1, ref).sort(-1)[0]
        assert torch.equal(dm, dr), f"stage {s}: kNN distance multisets differ"
        srt = d.sort(-1)[0]
        untied = srt[..., k - 1] < srt[..., k] if d.shape[-1] > k else torch.ones_like(srt[..., 0], dtype=torch.bool)
        same = (mine.sort(-1)[0] == ref.sort(-1)[0]).all(-1)
        assert bool(same[untied].all()), f"stage {s}: index sets differ where no tie exists"
        n_equal, n_total = n_equal + int(same.sum()), n_total + same.numel()
        if dt == torch.bfloat16:  # the kernel contract: the bf16 rounding points written out in fp32
            assert torch.equal(geo.knn_dist_bf16(new_xy, xy), d)
        xy = new_xy
    assert n_total > 0


def test_bf16_fps_distance_written_out():
    torch.manual_seed(0)
    xy = (torch.randint(0, 24, (3, 200, 2)).float() / 24).to(torch.bfloat16)
    c = xy[:, 17:18]
    assert torch.equal(geo.fps_dist_bf16(xy, c), torch.sum((xy - c) ** 2, -1).float())


def test_product_module_has_the_reference_parameter_tree():
    """medplib_b200.model.geo_sampler.GeoRegionSampler: constructor arguments, parameter names and shapes of the
    reference's module (the golden state dicts ARE the reference's), and no CPU path."""
    from medplib_b200 import _lib
    from medplib_b200.model.geo_sampler import GeoRegionSampler
    for case in CASES:
        d, out_dim, n_init, subs, neighs = case["cfg"]
        mod = GeoRegionSampler(input_dim=d, output_dim=out_dim, num_init_point=n_init, num_sub_point=subs,
                               num_neighbor=neighs, pooler_mode=case["pooler"])
        mine = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
        assert mine == {k: tuple(v.shape) for k, v in case["sd"].items()}
        mod.load_state_dict(case["sd"], strict=True)
    with pytest.raises(_lib.MplError):
        mod([torch.zeros(576, d)], [[torch.ones(24, 24)]], torch.float32, torch.float32)
    with pytest.raises(NotImplementedError):
        GeoRegionSampler(8, 8, 8, [4], [2], pooler_mode="sum")


def test_model_instantiates_the_sampler_only_when_configured():
    from medplib_b200.model import MedPLIBForCausalLM, MedPLIBMoELlamaConfig

    def make(**extra):
        cfg = MedPLIBMoELlamaConfig(hidden_size=16, intermediate_size=32, num_hidden_layers=1, num_attention_heads=1,
                                    num_key_value_heads=1, vocab_size=50, max_position_embeddings=64,
                                    mm_projector_type="mlp2x_gelu", max_sample_point=512)
        cfg.clip_config = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=1,
                               image_size=28, patch_size=14, layer_norm_eps=1e-5)
        cfg.sam_config = dict(image_size=256, embed_dim=64, depth=1, num_heads=1)
        cfg.moe = dict(num_experts=[2], top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0, min_capacity=0,
                       use_residual=False, router_aux_loss_coef=0.01, moe_layers_idx=None, moe_mode="dense", ep_size=1)
        for k, v in extra.items():
            setattr(cfg, k, v)
        return MedPLIBForCausalLM(cfg, test_only=True, seg_token_idx=42)

    assert not hasattr(make().get_model(), "region_geo_sampler")
    s = make(region_geo_sampler=True, sampler_pooler_mode="mean").get_model().region_geo_sampler
    assert (s.input_dim, s.output_dim, s.num_init_point, s.num_sub_point, s.num_neighbor, s.pooler_mode) == \
        (32, 16, 512, [128, 32], [24, 24], "mean")


def test_padded_gemm_operands_reproduce_the_grouped_linears():
    """Host logic of the product module (runs on CPU: pure tensor packing): the zero-padded operands `_weights()` builds
    for the point-table layout [features | row/H | col/W | 0...] give exactly what diff_projector / the 1x1 agg conv give
    on the unpadded rows (fp32 evaluation of the bf16 operands)."""
    from medplib_b200.model.geo_sampler import GeoRegionSampler
    torch.manual_seed(0)
    d = 40
    mod = GeoRegionSampler(d, 16, 32, [8, 4], [4, 2]).to(torch.bfloat16)
    ld = (d + 2 + 63) // 64 * 64
    rows = 11
    local = torch.randn(rows, d + 2).to(torch.bfloat16)
    anchor = torch.randn(rows, d + 2).to(torch.bfloat16)
    for s, (wd, bd, wa) in enumerate(mod._weights()):
        assert wd.shape == (ld, ld) and bd.shape == (ld,) and wa.shape == (d, 2 * ld)
        pad = lambda t: torch.cat([t, torch.zeros(rows, ld - (d + 2), dtype=t.dtype)], 1)
        dl, ag = mod.diff_projector_list[s], mod.agg_projector_list[s]
        diff = (local.float() - anchor.float()).to(torch.bfloat16)
        want_d = torch.nn.functional.linear(diff.float(), dl.weight.float(), dl.bias.float())
        got_d = torch.nn.functional.linear(pad(diff).float(), wd.float(), bd.float())
        assert torch.equal(got_d[:, :d + 2], want_d) and not bool(got_d[:, d + 2:].any())
        dproj = want_d.to(torch.bfloat16)
        cat = torch.cat([dproj, anchor], 1).float()                         # what the reference's conv sees
        want_a = torch.nn.functional.conv1d(cat.t()[None], ag.net[0].weight.float(), ag.net[0].bias.float())[0].t()
        got_a = torch.nn.functional.linear(torch.cat([pad(dproj), pad(anchor)], 1).float(), wa.float(),
                                           ag.net[0].bias.float())
        torch.testing.assert_close(got_a, want_a, rtol=1e-6, atol=1e-6)
    # the cache follows parameter updates
    before = mod._weights()[0][0].clone()
    with torch.no_grad():
        mod.diff_projector_list[0].weight.add_(1.0)
    assert not torch.equal(mod._weights()[0][0], before)
