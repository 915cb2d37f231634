"""Oracle: Ferret-style geometric region sampler (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Functional restatement of model/rp_sampler/GeoSampler.py: rand_sample_repeat :19-29,
point_sample :31-56, farthest_point_sample :59-80, square_distance :101-121, knn_point :124-136, ConvReLULN1D
:139-157 and GeoRegionSampler.forward :229-345, over a state dict with the reference's parameter names
(``diff_projector_list.{i}.*``, ``agg_projector_list.{i}.net.0.*`` / ``.norm.*``, ``flatten_projector.*``,
``dim_projector.*``).

Randomness: the reference draws from torch's GLOBAL CPU generator (``randperm`` / ``randint`` in rand_sample_repeat,
``randint`` for the FPS start). This restatement makes the same calls in the same order, so after the same
``torch.manual_seed`` it selects the same points as the reference; ``draws`` injects recorded draws instead.

Ties: coordinates are k/24 grid points (with repeats when a region has fewer than num_init_point pixels), so exact
distance ties are the rule, not the exception. FPS takes the FIRST maximum (what ``torch.max`` does on CPU; pinned to
the reference run). For kNN the reference's ``torch.topk(sorted=False)`` leaves the choice among equal distances at
the k-th place to the implementation (and on CUDA to the launch); the oracle — and the CUDA kernel — take the k
smallest by (distance, index), listed in that order. Pinning against the reference therefore checks (tests/
test_geo_cpu.py): sampled points and FPS indices equal, kNN distance multisets equal, and the whole forward equal to
the reference's output when the reference's own recorded kNN indices are injected (``knn_override``).
"""
import math

import torch
import torch.nn.functional as F


def rand_sample_repeat(x, max_len, draws=None):
    """GeoSampler.py:19-29. ``draws``: list consumed front to back with the index tensors the reference drew."""
    n = x.shape[0]
    if n < max_len:
        idx = draws.pop(0) if draws is not None else torch.randint(0, n, (max_len - n,))
        return torch.cat((x, x[idx]), dim=0)
    if n == max_len:
        return x
    idx = (draws.pop(0) if draws is not None else torch.randperm(n))[:max_len]
    return x[idx, :]


def sample_points(masks, num_init_point, draws=None):
    """:255-262 — normalised (row / H, col / W) of the non-zero pixels, resampled to num_init_point. fp32 [R, P, 2]."""
    hw = torch.tensor([masks[0].shape[0], masks[0].shape[1]])[None]
    return torch.stack([rand_sample_repeat(m.cpu().nonzero() / hw, num_init_point, draws) for m in masks])


def point_features(fmap, pos, original_dtype, return_dtype):
    """:263-276 — bilinear grid_sample (align_corners=True) in fp32 of the (x, y)-flipped, original_dtype-rounded
    coordinates. fmap [h*w, C]; pos [R, P, 2] -> [R, P, C] in return_dtype."""
    h = w = int(math.sqrt(fmap.shape[0]))
    c = fmap.shape[-1]
    f = fmap.reshape(h, w, c).permute(2, 0, 1).unsqueeze(0).repeat(pos.shape[0], 1, 1, 1).to(original_dtype)
    coords = pos.flip(dims=(2,)).type(original_dtype).unsqueeze(2)
    s = F.grid_sample(f.float(), (2.0 * coords - 1.0).float(), align_corners=True).to(return_dtype).squeeze(3)
    return s.transpose(-2, -1)


def fps(xyz, npoint, start):
    """:59-80 with the start index given. xyz [B, N, 2] -> long [B, npoint]; first maximum on ties."""
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long)
    distance = torch.ones(B, N) * 1e10
    farthest = start.clone()
    bi = torch.arange(B)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[bi, farthest, :].view(B, 1, 2)
        dist = torch.sum((xyz - centroid) ** 2, -1)
        distance = torch.min(distance, dist)
        farthest = torch.max(distance, -1)[1]
    return centroids


def square_distance(src, dst):
    """:101-121, same operation order (and therefore the same bf16 roundings)."""
    B, N, _ = src.shape
    M = dst.shape[1]
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def knn(nsample, xyz, new_xyz):
    """:124-136 with the tie rule of this module: the nsample smallest by (distance, index), in that order."""
    d = square_distance(new_xyz, xyz)
    return torch.sort(d.float(), dim=-1, stable=True)[1][..., :nsample]


def index_points(points, idx):
    """:83-98."""
    B = points.shape[0]
    bi = torch.arange(B).view(B, *([1] * (idx.dim() - 1))).expand_as(idx)
    return points[bi, idx, :]


def geo_region_sampler(sd, p, feature_map, region_masks, original_dtype, return_dtype, num_init_point, num_sub_point,
                       num_neighbor, pooler_mode="mean", draws=None, fps_start=None, knn_override=None, record=None):
    """GeoRegionSampler.forward :229-345. feature_map: list of [h*w, C]; region_masks: list of lists of [H, W] masks.
    Returns the list (per image) of [num_mask, output_dim] tensors or None.
    ``fps_start`` / ``knn_override``: per-stage lists replacing the random FPS start / the kNN choice;
    ``record``: dict that receives the intermediate points, FPS and kNN indices."""
    assert len(feature_map) == len(region_masks)
    pts, feas, img_ids = [], [], []
    for i, (fmap, masks) in enumerate(zip(feature_map, region_masks)):
        if len(masks) != 0:
            pos = sample_points(masks, num_init_point, draws)
            pts.append(pos)
            feas.append(point_features(fmap, pos, original_dtype, return_dtype))
            img_ids.extend([i] * len(pos))
    if not pts:
        return [None] * len(region_masks)
    xy = torch.cat(pts, 0).to(return_dtype)
    fea = torch.cat(feas, 0)
    if record is not None:
        record.update(points=xy.clone(), features=fea.clone(), fps=[], knn=[], stage_out=[])
    for s, (S, k) in enumerate(zip(num_sub_point, num_neighbor)):
        xy = xy.contiguous()
        start = fps_start[s] if fps_start is not None else torch.randint(0, xy.shape[1], (xy.shape[0],), dtype=torch.long)
        fi = fps(xy, S, start)
        new_xy, new_fea = index_points(xy, fi), index_points(fea, fi)
        idx = knn_override[s] if knn_override is not None else knn(k, xy, new_xy)
        local = torch.cat([index_points(fea, idx), index_points(xy, idx)], -1)  # [R, S, k, d+2]
        anchor = torch.cat([new_fea, new_xy], -1).unsqueeze(-2)
        diff = F.linear(local - anchor, sd[f"{p}diff_projector_list.{s}.weight"], sd[f"{p}diff_projector_list.{s}.bias"])
        g = torch.cat([diff, anchor.repeat(1, 1, k, 1)], -1)  # [R, S, k, 2(d+2)]
        y = F.relu(F.linear(g, sd[f"{p}agg_projector_list.{s}.net.0.weight"].squeeze(-1),
                            sd[f"{p}agg_projector_list.{s}.net.0.bias"]))
        y = F.layer_norm(y, (y.shape[-1],), sd[f"{p}agg_projector_list.{s}.norm.weight"],
                         sd[f"{p}agg_projector_list.{s}.norm.bias"], 1e-5)
        y = y.permute(0, 1, 3, 2).flatten(0, 1)  # [R*S, d, k]
        y = F.avg_pool1d(y, k) if pooler_mode == "mean" else F.adaptive_max_pool1d(y, 1)
        fea = y.reshape(xy.shape[0], S, -1)
        xy = new_xy
        if record is not None:
            record["fps"].append(fi)
            record["knn"].append(idx)
            record["stage_out"].append(fea.clone())
    x = F.linear(fea.flatten(1, -1), sd[p + "flatten_projector.weight"], sd[p + "flatten_projector.bias"])
    out = F.linear(x, sd[p + "dim_projector.weight"], sd[p + "dim_projector.bias"])
    ids = torch.tensor(img_ids)
    return [out[ids == i] if bool((ids == i).any()) else None for i in range(len(region_masks))]


# ------------------------------------------------------------------------------- explicit bf16 roundings (kernel contract)
def _r(x):
    return x.to(torch.bfloat16).float()


def fps_dist_bf16(xyz, c):
    """The bf16 rounding points of ``torch.sum((xyz - c) ** 2, -1)`` written out in fp32 (what geo_fps_kernel does)."""
    x, cc = xyz.float(), c.float()
    dx, dy = _r(x[..., 0] - cc[..., 0]), _r(x[..., 1] - cc[..., 1])
    return _r(_r(dx * dx) + _r(dy * dy))


def knn_dist_bf16(q, xyz):
    """The bf16 rounding points of square_distance(q, xyz) written out in fp32 (what geo_knn_kernel does).
    q [..., S, 2], xyz [..., N, 2] bf16-representable -> [..., S, N]."""
    q, x = q.float(), xyz.float()
    dot = _r(q[..., :, None, 0] * x[..., None, :, 0] + q[..., :, None, 1] * x[..., None, :, 1])
    sq = _r(_r(q[..., 0] * q[..., 0]) + _r(q[..., 1] * q[..., 1]))[..., :, None]
    sx = _r(_r(x[..., 0] * x[..., 0]) + _r(x[..., 1] * x[..., 1]))[..., None, :]
    return _r(_r(-2.0 * dot + sq) + sx)
