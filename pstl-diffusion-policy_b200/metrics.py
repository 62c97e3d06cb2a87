"""Evaluation metrics of the sampling test on the device.

Drop-ins for the reference's ``nusc_api.measure_diversity`` / ``measure_extra_diversity`` (nusc_api.py:817-936),
``utils.compute_entropy`` (utils.py:388-417) and ``nusc_train.compute_ade_fde`` (nusc_train.py:877-887).  Upstream
runs them on the host (numpy masked arrays, one scipy ConvexHull per (scene, mode, step), torch.histogramdd on CPU);
here the masked std and the hull areas are one CUDA kernel (``pstl_diversity``) and the rest are device tensor ops.
"""
import math

import numpy as np
import torch

from . import native as _nv


def _masked_lane_stats(per_lane, lane_valid):
    """(mean over valid (scene, lane) cells, per-scene mean over valid lanes, three per-lane columns with invalid -> 0)
    of a (bs, 3) array — the np.ma reductions at the end of measure_diversity (nusc_api.py:832-838, 870-875)."""
    v = lane_valid.to(per_lane.dtype)
    n = v.sum()
    overall = (per_lane * v).sum() / n if float(n) > 0 else per_lane.new_tensor(float("nan"))
    # np.mean of a fully masked row returns masked; its .data is 0 there
    rows = torch.where(v.sum(-1) > 0, (per_lane * v).sum(-1) / torch.clamp(v.sum(-1), min=1.0), torch.zeros_like(v.sum(-1)))
    cols = [per_lane[:, i] * v[:, i] for i in range(3)]
    return overall, rows, cols


def measure_diversity(trajs, scores, valids, nt):
    """trajs (bs, m, 3, nt*2), scores (bs, m, 3), valids (bs, m, 3) ->
    (ma_std_avg, ma_vol_avg_each, (std_overall (bs,), std0, std1, std2), (vol_overall (bs,), vol0, vol1, vol2)),
    scalars as numpy floats and the lists as numpy arrays, as upstream returns them."""
    _nv.require_cuda(trajs, "trajs")
    bs, m = trajs.shape[0], trajs.shape[1]
    t, s, v = _nv.f32(trajs.reshape(bs, m, 3, nt * 2)), _nv.f32(scores.reshape(bs, m, 3)), _nv.f32(valids.reshape(bs, m, 3))
    std = torch.empty((bs, 3), dtype=torch.float32, device=t.device)
    vol = torch.empty((bs, 3), dtype=torch.float32, device=t.device)
    _nv.check(_nv.lib().pstl_diversity(_nv.fptr(t), _nv.fptr(s), _nv.fptr(v), bs, m, nt, _nv.fptr(std), _nv.fptr(vol),
                                       _nv.stream()), "pstl_diversity")
    lane_valid = v[:, 0, :] != 0
    so, sr, sc = _masked_lane_stats(std, lane_valid)
    vo, vr, vc = _masked_lane_stats(vol, lane_valid)
    packed = torch.stack([so.reshape(1).expand(bs), sr, *sc, vo.reshape(1).expand(bs), vr, *vc], 0).cpu().numpy()  # one read-back
    return (np.float32(packed[0, 0]), np.float64(packed[5, 0]), (packed[1], packed[2], packed[3], packed[4]),
            (packed[6].astype(np.float64), packed[7].astype(np.float64), packed[8].astype(np.float64),
             packed[9].astype(np.float64)))


def compute_entropy(x, mask, n_bins=10, x_min=None, x_max=None):
    """histogram entropy (bits) of every row of x (N, m) over its unmasked entries (utils.py:388-417): n_bins equal bins
    on [min-1e-5, max+1e-5] of the row (or the given range), probabilities clipped at 1e-5 inside the log."""
    clip = 1e-5
    big = torch.full_like(x, float("inf"))
    x_hi = torch.where(mask == 0, big, x)      # masked entries fall outside every bin
    x_lo = torch.where(mask == 0, -big, x)
    if x_min is None:
        lo = x_hi.min(dim=1)[0] - clip
        hi = x_lo.max(dim=1)[0] + clip
    else:
        lo = torch.full_like(x[:, 0], float(x_min))
        hi = torch.full_like(x[:, 0], float(x_max))
    alphas = torch.linspace(0.0, 1.0, n_bins + 1, device=x.device)[None, :]
    edges = lo[:, None] * (1 - alphas) + hi[:, None] * alphas
    inside = (x_hi[:, :, None] >= edges[:, None, :-1]) & (x_hi[:, :, None] < edges[:, None, 1:])
    counts = inside.float().sum(dim=1)
    probs = counts / torch.clip(counts.sum(dim=-1, keepdim=True), clip)
    return (-probs * torch.log2(torch.clip(probs, clip))).sum(dim=-1)


def compute_area(x, y, th, val, bs, nt, m):
    """occupied area of the ego-frame positions on a 100x100 grid per (scene, mode) (nusc_api.py:880-894);
    torch.histogramdd semantics (equal bins over [min, max] of the data, last edge inclusive) without leaving the GPU."""
    val = val.reshape(bs * 3, m, nt, 1)
    x_rel = x * torch.cos(th) + y * torch.sin(th)
    y_rel = -x * torch.sin(th) + y * torch.cos(th)
    xy = (torch.stack([x_rel, y_rel], dim=-1) * val).reshape(bs * 3, m * nt, 2)
    lo, hi = xy.min(dim=1, keepdim=True)[0], xy.max(dim=1, keepdim=True)[0]
    # histogramdd widens a degenerate range by +-0.5
    same = hi == lo
    lo = torch.where(same, lo - 0.5, lo)
    hi = torch.where(same, hi + 0.5, hi)
    length = hi - lo
    idx = torch.clamp(torch.floor((xy - lo) / length * 100.0).long(), 0, 99)
    flat = idx[..., 0] * 100 + idx[..., 1]
    occ = torch.zeros((bs * 3, 10000), dtype=torch.float32, device=x.device)
    occ.scatter_(1, flat, 1.0)
    area = occ.mean(dim=1) * length[:, 0, 0] * length[:, 0, 1]
    return area.mean()


def measure_extra_diversity(trajs, scores, valids, nt, controls, wmin, wmax, amin, amax):
    """entropy of the scores / controls and occupied area over the accepted samples (nusc_api.py:897-936).
    trajs (bs, m, 3, nt*4), scores / valids (bs, m, 3), controls (bs, m, 3, nt*2) -> dict of 0-d tensors."""
    bs, m, _ = scores.shape
    trajs = trajs.permute(0, 2, 1, 3).reshape(bs * 3, m, nt, 4)
    scores = scores.permute(0, 2, 1).reshape(bs * 3, m)
    valids = valids.permute(0, 2, 1).reshape(bs * 3, m)
    controls = controls.permute(0, 2, 1, 3).reshape(bs * 3, m, nt, 2)
    valids = valids * (scores > 0).float()
    ent_s = compute_entropy(scores, valids)
    rev = lambda t: t.permute(0, 2, 1).reshape(bs * 3 * nt, m)
    valids_rev = valids[:, None].repeat(1, nt, 1).reshape(bs * 3 * nt, m)
    x_ = trajs[..., 0] - trajs[:, :, 0:1, 0]
    y_ = trajs[..., 1] - trajs[:, :, 0:1, 1]
    ent_w = compute_entropy(rev(controls[..., 0]), valids_rev, x_min=wmin, x_max=wmax)
    ent_a = compute_entropy(rev(controls[..., 1]), valids_rev, x_min=amin, x_max=amax)
    area = compute_area(x_, y_, trajs[..., 2], valids_rev, bs, nt, m)
    return {"ent_s": ent_s.mean(), "ent_w": ent_w.mean(), "ent_a": ent_a.mean(), "ent_wa": ent_w.mean() + ent_a.mean(),
            "area": area}


def compute_ade_fde(gt_trajs, est_trajs, mask):
    """min-over-samples average / final displacement (squared, all state columns) against the recorded trajectory
    (nusc_train.py:877-887); masked samples count 1e4 per column.  gt (bs, nt, k), est (bs, m, 3, nt, k) or (bs*m*3, nt, k)."""
    bs, nt, k = gt_trajs.shape
    mask = mask.reshape(bs, -1)[:, :, None, None]
    est = est_trajs.reshape(bs, -1, nt, k)
    err_t = torch.sum(torch.square((gt_trajs[:, None] - est) * mask + (1 - mask) * 10000), dim=-1)
    ade = torch.mean(torch.min(torch.mean(err_t, dim=-1), dim=-1)[0])
    fde = torch.mean(torch.min(err_t[:, :, -1], dim=-1)[0])
    return ade, fde
