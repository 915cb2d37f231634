"""GeoRegionSampler — the Ferret-style geometric region sampler behind ``--region_geo_sampler``.

Drop-in for model/rp_sampler/GeoSampler.py:162-345: same constructor arguments, module tree and parameter names
(``diff_projector_list.{i}``, ``agg_projector_list.{i}.net.0`` / ``.norm``, ``pooler_list``, ``flatten_projector``,
``dim_projector``), same ``forward(feature_map, region_masks, original_dtype, return_dtype)`` contract. The forward
runs hand-written sm_100a kernels (csrc/geo.cu: point sampling, farthest-point sampling, kNN, grouping, LayerNorm +
pooling) and the tcgen05 / streaming GEMMs for the four Linears; there is no CPU path.

Randomness is the reference's: ``rand_sample_repeat`` (:19-29) and the FPS start (:70) draw from torch's global CPU
generator with the same calls in the same order, so a seeded run picks the reference's points. Ties (exact on the 24x24
grid): FPS keeps the first maximum, kNN the k smallest by (distance, index) — the reference's ``topk(sorted=False)``
leaves that choice to the backend.
"""
import math

import torch
import torch.nn as nn

from .. import _lib, ops

bf16 = torch.bfloat16


class ConvReLULN1D(nn.Module):
    """Parameter holder with the reference's names (GeoSampler.py:139-157): ``net.0`` = Conv1d(k=1), ``norm``."""

    def __init__(self, in_channels, out_channels, kernel_size=1, bias=True):
        super().__init__()
        self.act = nn.ReLU(inplace=True)
        self.net = nn.Sequential(nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, bias=bias), self.act)
        self.norm = nn.LayerNorm(out_channels)

    def forward(self, x):
        raise _lib.MplError("ConvReLULN1D runs inside GeoRegionSampler.forward (fused GEMM epilogue + geo_ln_pool)")


def rand_sample_repeat(x, max_len):
    """GeoSampler.py:19-29 on the host (same generator calls as the reference)."""
    n = x.shape[0]
    if n < max_len:
        return torch.cat((x, x[torch.randint(0, n, (max_len - n,))]), dim=0)
    if n == max_len:
        return x
    return x[torch.randperm(n)[:max_len], :]


class GeoRegionSampler(nn.Module):
    def __init__(self, input_dim, output_dim, num_init_point, num_sub_point, num_neighbor, pooler_mode="mean"):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.num_init_point, self.num_sub_point, self.num_neighbor = num_init_point, num_sub_point, num_neighbor
        if pooler_mode not in ("mean", "max"):
            raise NotImplementedError(f"{pooler_mode} is not supported.")
        self.pooler_mode = pooler_mode
        self.diff_projector_list = nn.ModuleList()
        self.agg_projector_list = nn.ModuleList()
        self.pooler_list = nn.ModuleList()
        for ii in range(len(num_sub_point)):
            self.diff_projector_list.append(nn.Linear(input_dim + 2, input_dim + 2))
            self.agg_projector_list.append(ConvReLULN1D(2 * (input_dim + 2), input_dim))
            self.pooler_list.append(nn.AvgPool1d(kernel_size=num_neighbor[ii]) if pooler_mode == "mean"
                                    else nn.AdaptiveMaxPool1d(output_size=1))
        self.flatten_projector = nn.Linear(input_dim * num_sub_point[-1], input_dim)
        self.dim_projector = nn.Linear(input_dim, output_dim)
        self._packed = None
        self.trace = None  # set to a dict to receive the points / FPS / kNN indices of the next forward (tests)

    # -------------------------------------------------------------------------------------------- padded GEMM operands
    def _weights(self):
        """The grouped Linears as GEMM operands over the padded point-table rows (K = ld = round_up(d + 2, 64)):
        diff [ld, ld] (zero rows / columns in the padding), agg [d, 2 * ld] = [W[:, :d+2] | 0 | W[:, d+2:] | 0].
        Rebuilt when a parameter changes (version counters)."""
        ps = [q for m in (self.diff_projector_list, self.agg_projector_list) for q in m.parameters()]
        key = tuple((q.data_ptr(), q._version) for q in ps)
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        d = self.input_dim
        ld = (d + 2 + 63) // 64 * 64
        packed = []
        for dl, ag in zip(self.diff_projector_list, self.agg_projector_list):
            dev = dl.weight.device
            wd = torch.zeros((ld, ld), dtype=bf16, device=dev)
            wd[:d + 2, :d + 2] = dl.weight.detach()
            bd = torch.zeros((ld,), dtype=bf16, device=dev)
            bd[:d + 2] = dl.bias.detach()
            w = ag.net[0].weight.detach().squeeze(-1)
            wa = torch.zeros((d, 2 * ld), dtype=bf16, device=dev)
            wa[:, :d + 2] = w[:, :d + 2]
            wa[:, ld:ld + d + 2] = w[:, d + 2:]
            packed.append((wd, bd, wa))
        self._packed = (key, packed)
        return packed

    def forward(self, feature_map, region_masks, original_dtype, return_dtype):
        """feature_map: per image [h*w, C] (list or stacked tensor); region_masks: per image a list of [H, W] masks.
        Returns per image [num_mask, output_dim] or None (GeoSampler.py:229-345)."""
        assert len(feature_map) == len(region_masks)
        w0 = self.dim_projector.weight
        if not w0.is_cuda or w0.dtype != bf16 or original_dtype != bf16 or return_dtype != bf16:
            raise _lib.MplError("GeoRegionSampler: bf16 CUDA parameters / features only (medplib_b200 has no CPU path)")
        if torch.is_grad_enabled() and self.training and any(q.requires_grad for q in self.parameters()):
            raise _lib.MplError("GeoRegionSampler: the backward is not built (the reference never instantiates the module: "
                                "medplib_arch.py:134-142); freeze it or run under torch.no_grad()")
        dev = w0.device
        pts, img_of = [], []
        for i, masks in enumerate(region_masks):
            if len(masks) != 0:
                hw = torch.tensor([masks[0].shape[0], masks[0].shape[1]])[None]
                for m in masks:
                    pts.append(rand_sample_repeat(m.cpu().nonzero() / hw, self.num_init_point))
                    img_of.append(i)
        if not pts:
            return [None] * len(region_masks)
        fm = feature_map if torch.is_tensor(feature_map) else torch.stack(list(feature_map))
        h = w = int(math.sqrt(fm.shape[1]))
        d = self.input_dim
        ld = (d + 2 + 63) // 64 * 64
        R = len(pts)
        pts_dev = torch.stack(pts).float().to(dev)
        table = ops.geo_point_table(fm.contiguous(), torch.tensor(img_of, dtype=torch.int32, device=dev), pts_dev, h, w, ld)
        if self.trace is not None:
            self.trace.update(points=pts_dev, table=table, fps=[], knn=[], stage_out=[])
        n_stage = len(self.num_sub_point)
        for s, (S, k, (wd, bd, wa)) in enumerate(zip(self.num_sub_point, self.num_neighbor, self._weights())):
            N = table.shape[1]
            start = torch.randint(0, N, (R,), dtype=torch.long).to(device=dev, dtype=torch.int32)
            fi = ops.geo_fps(table, d, S, start)
            ki = ops.geo_knn(table, d, fi, k)
            a1, a2 = ops.geo_group(table, fi, ki)
            ops.linear(a1, wd, bias=bd, out=a2[:, :ld])  # diff_projector, written beside the anchor rows
            ag = self.agg_projector_list[s]
            y = ops.linear(a2, wa, bias=ag.net[0].bias, act="relu")
            last = s == n_stage - 1
            table = ops.geo_ln_pool(y, R, S, k, ag.norm.weight, ag.norm.bias, ag.norm.eps, self.pooler_mode,
                                    table=None if last else table, d=d, fps_idx=fi, ldo=d if last else ld)
            if self.trace is not None:
                self.trace["fps"].append(fi)
                self.trace["knn"].append(ki)
                self.trace["stage_out"].append(table)
        x = ops.linear(table.reshape(R, -1), self.flatten_projector.weight, bias=self.flatten_projector.bias)
        out = ops.linear(x, self.dim_projector.weight, bias=self.dim_projector.bias)
        ids = torch.tensor(img_of)
        return [out[torch.nonzero(ids == i).flatten().to(dev)] if bool((ids == i).any()) else None
                for i in range(len(region_masks))]
