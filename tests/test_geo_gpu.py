"""GPU parity of the GeoRegionSampler path (csrc/geo.cu + medplib_b200/model/geo_sampler.py) against
oracle/geo.py (pinned to the reference's module by tests/test_geo_cpu.py) and the reference's own bf16 run
(tests/golden/geo.pt). Bit-exact: sampled points, FPS indices, kNN indices (the (distance, index) rule). Floating point:
features / outputs within the stated multiples of a bf16 ulp of the output scale."""
import contextlib
import os

import pytest
import torch

from oracle import geo
from parity import BF16_ULP, close as _close

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
HERE = os.path.dirname(os.path.abspath(__file__))


@contextlib.contextmanager
def replay(draws):
    """torch.randint / torch.randperm return the recorded draws (the reference's, or the oracle's) front to back."""
    q = list(draws)
    o_ri, o_rp = torch.randint, torch.randperm
    torch.randint = lambda *a, **k: q.pop(0).clone()
    torch.randperm = lambda *a, **k: q.pop(0).clone()
    try:
        yield q
    finally:
        torch.randint, torch.randperm = o_ri, o_rp


def _blob(g, size, n_pix):
    m = torch.zeros(size, size, dtype=torch.long)
    x, y = int(torch.randint(size, (1,), generator=g)), int(torch.randint(size, (1,), generator=g))
    m[x, y] = 1
    for _ in range(n_pix):
        nb = [(x + dx, y + dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1)
              if (dx or dy) and 0 <= x + dx < size and 0 <= y + dy < size and m[x + dx, y + dy] == 0]
        if nb:
            x, y = nb[int(torch.randint(len(nb), (1,), generator=g))]
            m[x, y] = 1
    return m


def _module(sd, cfg, pooler, dev):
    from medplib_b200.model.geo_sampler import GeoRegionSampler
    d, out_dim, n_init, subs, neighs = cfg
    mod = GeoRegionSampler(d, out_dim, n_init, subs, neighs, pooler_mode=pooler)
    mod.load_state_dict(sd, strict=True)
    return mod.to(device=dev, dtype=bf16).eval()


def _run_both(sd, cfg, pooler, fmaps, masks, draws, dev, tol_ulps, tag):
    d, out_dim, n_init, subs, neighs = cfg
    n_stage = len(subs)
    mod = _module(sd, cfg, pooler, dev)
    mod.trace = {}
    with replay(draws) as left, torch.no_grad():
        out = mod(torch.stack(fmaps).to(dev), [[m.to(dev) for m in per] for per in masks], bf16, bf16)
    assert not left, "the module made fewer RNG calls than the reference"
    rec = {}
    ref = geo.geo_region_sampler(sd, "", fmaps, masks, bf16, bf16, n_init, subs, neighs, pooler,
                                 draws=list(draws[:len(draws) - n_stage]), fps_start=draws[len(draws) - n_stage:],
                                 record=rec)
    tr = mod.trace
    tab = tr["table"].cpu()
    assert torch.equal(tab[..., d:d + 2], rec["points"]), "point coordinates"
    assert not bool(tab[..., d + 2:].any()), "table padding"
    _close(tab[..., :d], rec["features"], 0.0, f"{tag} point features")  # same rounding points: bit-exact
    for s in range(n_stage):
        assert torch.equal(tr["fps"][s].cpu().long(), rec["fps"][s]), f"FPS indices, stage {s}"
        assert torch.equal(tr["knn"][s].cpu().long(), rec["knn"][s]), f"kNN indices, stage {s}"
        _close(tr["stage_out"][s].cpu()[..., :d], rec["stage_out"][s], tol_ulps * BF16_ULP, f"{tag} stage {s} features")
    for o, r in zip(out, ref):
        assert (o is None) == (r is None)
        if o is not None:
            assert o.shape == r.shape and o.dtype == bf16
            _close(o, r, tol_ulps * BF16_ULP, f"{tag} region features")
    return out, rec


def test_geo_sampler_vs_reference_bf16_run(dev):
    """The reference's own bf16 run (golden case 3): its RNG draws replayed; points and FPS indices equal ITS record,
    kNN a valid top-k of its distances, everything else against the oracle."""
    case = [c for c in torch.load(os.path.join(HERE, "golden", "geo.pt"), weights_only=False)
            if c["dtype"] == "torch.bfloat16"][0]
    sd = {k: v.to(bf16) for k, v in case["sd"].items()}
    fmaps = [f.to(bf16) for f in case["fmaps"]]
    masks = [[m.long() for m in per] for per in case["masks"]]
    draws = [t for _, t in case["draws"]]
    out, rec = _run_both(sd, tuple(case["cfg"]), case["pooler"], fmaps, masks, draws, dev, 1.0, "golden")
    for s in range(len(case["fps"])):
        assert torch.equal(rec["fps"][s], case["fps"][s])
    # the reference's output itself used another (equally valid) choice among tied neighbours: same scale, not equal
    for o, r in zip(out, case["out"]):
        if r is not None:
            assert o.shape == r.shape


@pytest.mark.parametrize("pooler,d,out_dim,n_init,subs,neighs,n_pix", [
    ("max", 1024, 4096, 512, [128, 32], [24, 24], [[40, 300], [], [575]]),   # the configuration of medplib_arch.py:136-141
    ("mean", 1024, 4096, 512, [128, 32], [24, 24], [[120]]),
    ("mean", 64, 80, 100, [50, 30], [20, 10], [[300], [520, 7]]),             # the reference unit test's point counts
])
def test_geo_sampler_matches_oracle(dev, pooler, d, out_dim, n_init, subs, neighs, n_pix):
    g = torch.Generator().manual_seed(11)
    from medplib_b200.model.geo_sampler import GeoRegionSampler
    proto = GeoRegionSampler(d, out_dim, n_init, subs, neighs, pooler_mode=pooler)
    sd = {}
    for k, v in proto.state_dict().items():
        if k.endswith("norm.weight"):
            sd[k] = (1.0 + 0.1 * torch.randn(v.shape, generator=g)).to(bf16)
        elif k.endswith("bias"):
            sd[k] = (0.05 * torch.randn(v.shape, generator=g)).to(bf16)
        else:
            sd[k] = (torch.randn(v.shape, generator=g) / (v.shape[1] ** 0.5)).to(bf16)
    fmaps = [(0.5 * torch.randn(576, d, generator=g)).to(bf16) for _ in n_pix]
    masks = [[_blob(g, 24, n) for n in per] for per in n_pix]
    # record the draws a plain seeded run makes (oracle order == reference order == module order)
    draws = []
    o_ri, o_rp = torch.randint, torch.randperm

    def ri(*a, **k):
        draws.append(o_ri(*a, **k))
        return draws[-1]

    def rp(*a, **k):
        draws.append(o_rp(*a, **k))
        return draws[-1]

    torch.manual_seed(5)
    torch.randint, torch.randperm = ri, rp
    try:
        for per in masks:
            if per:
                geo.sample_points(per, n_init)
        R = sum(len(per) for per in masks)
        N = n_init
        for S in subs:
            torch.randint(0, N, (R,), dtype=torch.long)
            N = S
    finally:
        torch.randint, torch.randperm = o_ri, o_rp
    _run_both(sd, (d, out_dim, n_init, subs, neighs), pooler, fmaps, masks, draws, dev, 4.0, f"d={d} {pooler}")


def test_geo_sampler_no_regions_and_refusals(dev):
    from medplib_b200 import _lib
    from medplib_b200.model.geo_sampler import GeoRegionSampler
    mod = GeoRegionSampler(64, 80, 32, [8, 4], [4, 2]).to(device=dev, dtype=bf16).eval()
    fm = torch.zeros(2, 576, 64, dtype=bf16, device=dev)
    assert mod(fm, [[], []], bf16, bf16) == [None, None]
    with pytest.raises(_lib.MplError):
        mod(fm, [[torch.ones(24, 24)], []], torch.float32, torch.float32)
    mod.train()
    with pytest.raises(_lib.MplError):
        mod(fm, [[torch.ones(24, 24)], []], bf16, bf16)


def test_model_region_path_uses_the_geo_sampler(dev):
    """config.region_geo_sampler: the region slot of the spliced prompt holds GeoRegionSampler(raw CLIP features)
    (medplib_arch.py:204-208, 229, 285-289) instead of the adapter's sample mean."""
    from medplib_b200.model import MedPLIBForCausalLM, MedPLIBMoELlamaConfig
    from test_model_gpu import CLIP_CFG
    torch.manual_seed(0)
    cfg = MedPLIBMoELlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=1, num_attention_heads=2,
                                num_key_value_heads=2, vocab_size=300, rms_norm_eps=1e-5, max_position_embeddings=512,
                                mm_vision_select_layer=-2, mm_projector_type="mlp2x_gelu", max_sample_point=64,
                                initializer_range=0.06)
    cfg.clip_config = CLIP_CFG
    cfg.sam_config = dict(image_size=256, embed_dim=128, depth=1, num_heads=2)
    cfg.moe = dict(num_experts=[2], top_k_experts=1, capacity_factor=1.5, eval_capacity_factor=2.0, min_capacity=0,
                   use_residual=False, router_aux_loss_coef=0.01, moe_layers_idx=None, moe_mode="dense", ep_size=1)
    cfg.region_geo_sampler, cfg.sampler_pooler_mode = True, "max"
    m = MedPLIBForCausalLM(cfg, test_only=True, seg_token_idx=299, use_mm_start_end=True)
    m.config.mm_use_im_start_end = True
    m = m.to(bf16).to(dev).eval()
    sampler = m.get_model().region_geo_sampler
    assert sampler.num_init_point == 64 and sampler.num_sub_point == [128, 32] and sampler.pooler_mode == "max"
    sampler.num_sub_point, sampler.num_neighbor = [32, 32], [8, 8]  # 64 initial points: keep S <= N
    sampler.flatten_projector = torch.nn.Linear(128 * 32, 128).to(device=dev, dtype=bf16)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(3, 290, (1, 12), generator=g)
    ids[0, 2], ids[0, 6] = -200, -300
    clip_img = torch.randn(1, 3, 56, 56, generator=g).to(bf16).to(dev)
    mask = _blob(g, 24, 40).to(dev)
    am = torch.ones_like(ids, dtype=torch.bool)
    with torch.no_grad():
        torch.manual_seed(9)
        _, _, _, emb, _ = m.prepare_inputs_labels_for_multimodal(ids.to(dev), am.to(dev), None, None, clip_img,
                                                                 [[mask]], [[True]])
        torch.manual_seed(9)
        raw = m.get_vision_tower()(clip_img)
        want = sampler(raw, [[mask]], bf16, bf16)[0]
    # the spliced row of the region token: 1 (bos..) the prompt keeps its order; find the row equal to `want`
    hit = (emb[0].float() - want[0].float()).abs().amax(-1) == 0
    assert int(hit.sum()) == 1, "exactly one spliced row carries the geo-sampled region feature"
