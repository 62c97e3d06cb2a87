"""Hot-path functions of the reference's ``nusc_train.py`` with the same names and signatures,
running on the CUDA kernels of libpstl_b200.so (reference nusc_train.py line ranges in each docstring).

Two calling styles are served by the same functions:
  * reference style — dense per-row dicts exactly as upstream builds them (``rows_per_scene = 1``);
  * scene-indexed — ``augment_batch_data`` / ``pre_prepare_stl_cache`` return lazy dicts that carry a
    ``ScenePack`` (compact scene tensors + per-row state/pSTL/mode), so the kernels index scenes
    instead of reading K-times-replicated copies; the dense tensors are only materialised if a
    caller actually reads them.
Beyond the sampling test: ``trajopt`` (traj-opt data generation), ``compute_policy_loss``,
``train_step_ddpm`` / ``train_step_rect`` / ``run_training`` (the README's two training stages).
Out of scope: dataset / NuScenes access, pSTL calibration, visualisation, the VAE / BC baselines.
"""
import argparse
import os
import math
import time

import numpy as np
import torch

from . import native as _nv
from .stl_d_lib import *  # noqa: F401,F403  (the reference does the same star import)
from .stl_d_lib import AP, Always, And, Eventually, ListAnd, compile_formula, get_program

I_VAL = 0
I_X, I_Y, I_TH, I_V = 0, 1, 2, 3
I_VMIN, I_VMAX, I_DMIN, I_DMAX, I_DSAFE, I_THMAX = 0, 1, 2, 3, 4, 5


def dup(x, m):
    """(N,d) -> (N*m,d) (reference :20-21)."""
    return x.unsqueeze(1).repeat((1, m) + tuple(1 for _ in x.shape[1:])).reshape((-1,) + x.shape[1:])


def mul_n(x, n):
    """reference :253-256."""
    return x[:, None].repeat(1, n, *[1] * (x.dim() - 1)).reshape(x.shape[0] * n, *x.shape[1:])


def mask_mean(loss, mask, dim=None):
    """reference :23-27."""
    if dim is not None:
        return torch.mean(loss * mask, dim=dim) / torch.clip(torch.mean(mask, dim=dim), 1e-2)
    return torch.mean(loss * mask) / torch.clip(torch.mean(mask), 1e-2)


def dynamics(s, u):
    """reference :29-37 (one Euler derivative; elementwise, stays in PyTorch)."""
    th, v = s[..., 2], s[..., 3]
    return torch.stack([v * torch.cos(th), v * torch.sin(th), u[..., 0], u[..., 1]], dim=-1)


class _Rollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, us, dt):
        N, T = us.shape[0], us.shape[1]
        traj = torch.empty((N, T + 1, 4), dtype=torch.float32, device=us.device)
        _nv.check(_nv.lib().pstl_rollout(_nv.fptr(s), _nv.fptr(us), N, T, _nv.C.c_float(dt), _nv.fptr(traj),
                                         _nv.stream()), "pstl_rollout")
        ctx.save_for_backward(traj)
        ctx.dt = dt
        return traj

    @staticmethod
    def backward(ctx, g):
        (traj,) = ctx.saved_tensors
        N, T = traj.shape[0], traj.shape[1] - 1
        gs = torch.empty((N, 4), dtype=torch.float32, device=traj.device)
        gu = torch.empty((N, T, 2), dtype=torch.float32, device=traj.device)
        _nv.check(_nv.lib().pstl_rollout_bwd(_nv.fptr(traj), _nv.fptr(_nv.f32(g)), N, T, _nv.C.c_float(ctx.dt),
                                             _nv.fptr(gs), _nv.fptr(gu), _nv.stream()), "pstl_rollout_bwd")
        return gs, gu, None


def generate_trajs(s, us, dt):
    """(...,4) x (...,T,2) -> (...,T+1,4) Euler unicycle rollout (reference :39-49)."""
    assert s.shape[-1] == 4
    assert us.shape[-1] == 2
    assert us.shape[:-2] == s.shape[:-1]
    _nv.require_cuda(us, "controls")
    lead = us.shape[:-2]
    T = us.shape[-2]
    s2 = s.reshape(-1, 4).to(torch.float32).contiguous()
    u2 = us.reshape(-1, T, 2).to(torch.float32).contiguous()
    return _Rollout.apply(s2, u2, float(dt)).reshape(*lead, T + 1, 4)


def get_neighbor_trajs(neighbors, nt, dt, full=False):
    """constant-velocity neighbour prediction (reference :51-60)."""
    no_cmd = torch.zeros_like(neighbors[..., :2]).unsqueeze(-2).repeat(1, 1, nt - 1, 1)
    trajs = generate_trajs(neighbors[..., 1:5], no_cmd, dt)
    valids = neighbors[..., 0:1].unsqueeze(-2).repeat(1, 1, nt, 1)
    if full:
        lws = neighbors[..., 5:7].unsqueeze(-2).repeat(1, 1, nt, 1)
        return torch.cat([valids, trajs, lws], dim=-1)
    return torch.cat([valids, trajs], dim=-1)


# ---------------------------------------------------------------------------------------
# scene-indexed containers
# ---------------------------------------------------------------------------------------

_mode_cache = {}


def _mode_column(n_groups, device):
    """(n_groups*3, 1) float column 0,1,2,0,1,2,... (highlevel_dense, reference :753); cached per shape"""
    key = (int(n_groups), str(device))
    t = _mode_cache.get(key)
    if t is None:
        t = torch.tensor([0.0, 1.0, 2.0]).repeat(n_groups).reshape(-1, 1).to(device)
        if len(_mode_cache) > 8:
            _mode_cache.clear()
        _mode_cache[key] = t
    return t


class ScenePack:
    """Compact inputs of the scoring / guidance kernels for N = bs*S*3 chains
    (flat index n = (scene*S + sample)*3 + mode)."""

    def __init__(self, neighbors, lanes, state0, stlp, mode, valid, rows_per_scene):
        self.neighbors = _nv.f32(neighbors)
        self.lanes = [_nv.f32(l) for l in lanes]
        self.state0 = _nv.f32(state0)
        self.stlp = _nv.f32(stlp)
        self.mode = _nv.f32(mode)
        self.valid = _nv.f32(valid)
        self.rows_per_scene = int(rows_per_scene)
        self.N = self.state0.shape[0]

    @property
    def T(self):
        return self.neighbors.shape[2]

    def view(self):
        return _nv.make_scene_view(self.neighbors, self.lanes, self.rows_per_scene)

    @staticmethod
    def from_batch(batch, stlp_dense, S):
        bs = batch["currlane_wpts"].shape[0]
        m = S * 3
        nei = batch["neighbor_trajs_aug"] if "neighbor_trajs_aug" in batch else batch["neighbors_traj"][..., :7]
        state0 = dup(batch["ego_traj"][:, 0, :4], m)
        valids = torch.cat([batch["curr_id"], batch["left_id"], batch["right_id"]], dim=-1)
        valid = dup(valids, S).reshape(-1)
        mode = _mode_column(bs * S, nei.device).reshape(-1)
        return ScenePack(nei, [batch["currlane_wpts"], batch["leftlane_wpts"], batch["rightlane_wpts"]], state0,
                         stlp_dense.reshape(bs * m, 6), mode, valid, m)


class LazyBatch(dict):
    """dict whose dense (row-replicated) tensors are built only when somebody reads them."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._lazy = {}

    def set_lazy(self, key, fn):
        self._lazy[key] = fn

    def __missing__(self, key):
        if key in self._lazy:
            v = self._lazy.pop(key)()
            self[key] = v
            return v
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def keys_all(self):
        return list(dict.keys(self)) + list(self._lazy)


# ---------------------------------------------------------------------------------------
# STL spec (reference :95-140) with typed predicate leaves
# ---------------------------------------------------------------------------------------

def build_stl_cache(args):
    """[stl_curr, stl_left, stl_right] (reference :95-140).  Leaves are ``AP.predicate`` objects:
    the lambdas are the reference's expressions (generic path), the typed part feeds the fused kernel."""
    nt = args.nt
    norm = bool(getattr(args, "norm_stl", False))
    dv = _nv.DEN_VFACTOR if norm else _nv.DEN_ONE
    dd = _nv.DEN_DFACTOR if norm else _nv.DEN_ONE
    ds = _nv.DEN_SFACTOR if norm else _nv.DEN_ONE
    P = AP.predicate
    f = (lambda k: (lambda x: x[k])) if norm else (lambda k: (lambda x: 1.0))
    vf, df, sf = f("v_factor"), f("d_factor"), f("safe_factor")
    G = lambda ap: Always(0, nt, ap)
    keep_v_min = G(P(lambda x: (x["ego_traj"][..., I_V] - x["stlp"][..., I_VMIN]) / vf(x), _nv.SIG_V, 0, I_VMIN, 1, dv))
    keep_v_max = G(P(lambda x: (-x["ego_traj"][..., I_V] + x["stlp"][..., I_VMAX]) / vf(x), _nv.SIG_V, 1, I_VMAX, 0, dv))
    keep_d_min = G(P(lambda x: (x["x2curr_d"] - x["stlp"][..., I_DMIN]) / df(x), _nv.SIG_D_CURR, 0, I_DMIN, 1, dd))
    keep_d_max = G(P(lambda x: (-x["x2curr_d"] + x["stlp"][..., I_DMAX]) / df(x), _nv.SIG_D_CURR, 1, I_DMAX, 0, dd))

    def reach(side, sd, sa):
        band = And(P(lambda x: (x["x2%s_d" % side] - x["stlp"][..., I_DMIN]) / df(x), sd, 0, I_DMIN, 1, dd),
                   P(lambda x: (-x["x2%s_d" % side] + x["stlp"][..., I_DMAX]) / df(x), sd, 1, I_DMAX, 0, dd))
        rd = Eventually(0, nt // 2, G(band))
        rt = Eventually(0, nt // 2, G(P(lambda x: (x["stlp"][..., I_THMAX] - x["x2%s_th" % side]) / x["stlp"][..., I_THMAX],
                                        sa, 1, I_THMAX, 0, _nv.DEN_THMAX)))
        return rd, rt

    reach_left_d, reach_left_th = reach("left", _nv.SIG_D_LEFT, _nv.SIG_TH_LEFT)
    reach_right_d, reach_right_th = reach("right", _nv.SIG_D_RIGHT, _nv.SIG_TH_RIGHT)
    safe_list = [G(P(lambda x: (x["min_nei_d"] - x["stlp"][..., I_DSAFE]) / sf(x), _nv.SIG_NEI, 0, I_DSAFE, 1, ds))]
    keep_th_max = G(P(lambda x: (x["stlp"][..., I_THMAX] - x["x2curr_th"]) / x["stlp"][..., I_THMAX],
                      _nv.SIG_TH_CURR, 1, I_THMAX, 0, _nv.DEN_THMAX))
    stl_curr = ListAnd([keep_v_min, keep_v_max, keep_d_min, keep_d_max, keep_th_max] + safe_list)
    stl_left = ListAnd([keep_v_min, keep_v_max, reach_left_d, reach_left_th] + safe_list)
    stl_right = ListAnd([keep_v_min, keep_v_max, reach_right_d, reach_right_th] + safe_list)
    return [stl_curr, stl_left, stl_right]


def _fused_programs(stls_cac, T):
    """Compile the three formulas for the fused kernel; None if a leaf is an untyped lambda.  Nothing is cached on the
    formula objects: ``get_program`` keys the device programs on WHAT a formula compiles to (18 us for the driving spec),
    so a list that shares its first formula with an earlier one, or a ListAnd edited after its first use, never sees
    another list's programs."""
    try:
        return [get_program(compile_formula(f, fused=True)[0], 0, T, 1) for f in stls_cac]
    except ValueError:
        return None


def _spec(args, w_scale=1.0, a_scale=1.0, clip_controls=0):
    return _nv.make_spec(args.dt, args.smoothing_factor, args.ego_L, args.ego_W, w_scale, a_scale, clip_controls,
                         _lane_flags(args), 0)


def _lane_flags(args):
    """lane flag word of the kernels: bit 0 --clip_dist, bit 1 --inline (reference nusc_api.py:716-724, 732-733)"""
    return int(bool(args.clip_dist)) | (int(bool(getattr(args, "inline", False))) << 1)


def _default_anchors(args):
    """the fused scoring kernels hold the (refined_nL, refined_nW) = (4, 1) anchor grid of the reference's defaults in
    registers; any other grid is served by prep_stl_cache (pstl_car_distances) + the generic formula kernels"""
    return int(getattr(args, "refined_nL", 4)) == 4 and int(getattr(args, "refined_nW", 1)) == 1


def _check_fused_supported(args, what):
    if not _default_anchors(args):
        raise NotImplementedError("%s runs on the fused kernels, which are built for --refined_nL 4 --refined_nW 1; "
                                  "compute_stl_dense / prep_stl_cache take any anchor grid" % what)


class _CarDistances(torch.autograd.Function):
    """(min_dist, rad_sum) (N,K,T) of dist_between_two_cars(full=True) for any anchor grid (reference utils.py:465-526);
    differentiable w.r.t. the ego poses."""

    @staticmethod
    def forward(ctx, ego, sv_holder, args):
        nei, lanes, rps = sv_holder
        N, T = ego.shape[0], ego.shape[1]
        K = nei.shape[1]
        e = _nv.f32(ego)
        md = torch.empty((N, K, T), dtype=torch.float32, device=ego.device)
        rs = torch.empty((N, K, T), dtype=torch.float32, device=ego.device)
        part = torch.empty((N, K, T, 3), dtype=torch.float32, device=ego.device) if ctx.needs_input_grad[0] else None
        sv = _nv.make_scene_view(nei, lanes, rps)
        _nv.check(_nv.lib().pstl_car_distances(_nv.C.byref(sv), _nv.C.c_float(args.ego_L), _nv.C.c_float(args.ego_W),
                                               int(args.refined_nL), int(args.refined_nW), _nv.fptr(e), e.shape[2], N,
                                               _nv.fptr(md), _nv.fptr(rs), _nv.fptr(part), _nv.stream()),
                  "pstl_car_distances")
        ctx.part = part
        ctx.width = ego.shape[2]
        ctx.mark_non_differentiable(rs)
        return md, rs

    @staticmethod
    def backward(ctx, g_md, _g_rs):
        N, K, T, _ = ctx.part.shape
        ge = torch.zeros((N, T, ctx.width), dtype=torch.float32, device=g_md.device)
        ge[..., :3] = (g_md.unsqueeze(-1) * ctx.part).sum(dim=1)
        return ge, None, None


class _Predicates(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ego, sv_holder, args):
        nei, lanes, rps = sv_holder
        N, T = ego.shape[0], ego.shape[1]
        e = _nv.f32(ego)
        sig = torch.empty((N, 7, T), dtype=torch.float32, device=ego.device)
        part = torch.empty((N, 12, T), dtype=torch.float32, device=ego.device)
        sv = _nv.make_scene_view(nei, lanes, rps)
        _nv.check(_nv.lib().pstl_predicates(_nv.C.byref(sv), _nv.C.c_float(args.ego_L), _nv.C.c_float(args.ego_W),
                                            _lane_flags(args), _nv.fptr(e), e.shape[2], N, _nv.fptr(sig),
                                            _nv.fptr(part), _nv.stream()), "pstl_predicates")
        ctx.save_for_backward(part)
        ctx.width = ego.shape[2]
        return sig

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        N, _, T = p.shape
        ge = torch.zeros((N, T, ctx.width), dtype=torch.float32, device=p.device)
        for l in range(3):
            ge[..., 0] += g[:, 2 * l] * p[:, 3 * l]
            ge[..., 1] += g[:, 2 * l] * p[:, 3 * l + 1]
            ge[..., 2] += g[:, 2 * l + 1] * p[:, 3 * l + 2]
        ge[..., 0] += g[:, 6] * p[:, 9]
        ge[..., 1] += g[:, 6] * p[:, 10]
        ge[..., 2] += g[:, 6] * p[:, 11]
        return ge, None, None


def prep_stl_cache(x, args):
    """adds x2{curr,left,right}_{d,th} and min_nei_d to the dense dict (reference :74-93)."""
    ego = x["ego_traj"]
    _nv.require_cuda(ego, "ego_traj")
    pack = x.get("_pstl_pack") if isinstance(x, dict) else None
    if pack is not None:
        rep = x.get("_pstl_repeat", 1)
        if rep != 1:
            raise NotImplementedError("prep_stl_cache on candidate-stacked lazy inputs: use compute_stl_dense")
        holder = (pack.neighbors, pack.lanes, pack.rows_per_scene)
    else:
        holder = (_nv.f32(x["neighbors"]), [_nv.f32(x["%slane_wpts" % k]) for k in ("curr", "left", "right")], 1)
    sig = _Predicates.apply(ego, holder, args)
    for l, k in enumerate(("curr", "left", "right")):
        x["x2%s_d" % k], x["x2%s_th" % k] = sig[:, 2 * l], sig[:, 2 * l + 1]
    x["min_nei_d"] = sig[:, 6]
    full = getattr(args, "collision_loss", None) is not None
    if full or not _default_anchors(args):
        # reference :81-86, 142-148 on the general anchor-grid kernel
        md, rs = _CarDistances.apply(ego, holder, args)
        nei = holder[0]
        ind = nei[..., 0] if holder[2] == 1 else nei[..., 0][torch.arange(ego.shape[0], device=ego.device) // holder[2]]
        x["min_nei_d"] = torch.min(torch.clip(md - rs, -5, 20) * ind + (1 - ind) * 100, dim=1)[0]
        if full:
            x["min_centroid_d"], x["radius_sum"] = md * ind + (1 - ind) * 100, rs
    if getattr(args, "norm_stl", False):
        x["v_factor"] = torch.clip((x["stlp"][..., I_VMAX] - x["stlp"][..., I_VMIN]), 0.3)
        x["d_factor"] = torch.clip((x["stlp"][..., I_DMAX] - x["stlp"][..., I_DMIN]) * 5, 0.3)
        x["safe_factor"] = torch.clip(x["stlp"][..., I_DSAFE], 0.3)
    return x


def get_stl_scores(scores_list, stl_i):
    """reference :150-151."""
    return sum(scores_list[k] * (stl_i == k).float() for k in range(4))


def pre_prepare_stl_cache(batch_cuda, dense_trajs=None, detach=False, repeat_n=None, mono=False, mono_n=None,
                          gt_stlp=None):
    """reference :258-285.  On a batch produced by this module's ``augment_batch_data`` the result
    stays scene-indexed (lazy); on a plain dict it replicates exactly like upstream."""
    pack = batch_cuda.get("_pstl_pack") if not mono else None
    if pack is not None:
        out = LazyBatch()
        out["_pstl_pack"] = pack
        out["_pstl_repeat"] = 1 if repeat_n is None else int(repeat_n)
        rep = (lambda v: v) if repeat_n is None else (lambda v: v.repeat(repeat_n, *[1] * (v.dim() - 1)))
        for k_out, k_in in (("neighbors", "neighbors_dense"), ("currlane_wpts", "currlane_wpts_dense"),
                            ("leftlane_wpts", "leftlane_wpts_dense"), ("rightlane_wpts", "rightlane_wpts_dense")):
            out.set_lazy(k_out, (lambda ki: (lambda: rep(batch_cuda[ki])))(k_in))
        out["stlp"] = rep(batch_cuda["stlp_dense"])
        out["dense_valids"] = rep(batch_cuda["valids_dense"])
        out["gt_high_level"] = rep(batch_cuda["gt_high_level"])
        if dense_trajs is not None:
            out["ego_traj"] = dense_trajs
        return out
    if mono:
        stl_input = {
            "neighbors": mul_n(batch_cuda["neighbors_traj"], mono_n),
            "currlane_wpts": mul_n(batch_cuda["currlane_wpts"], mono_n),
            "leftlane_wpts": mul_n(batch_cuda["leftlane_wpts"], mono_n),
            "rightlane_wpts": mul_n(batch_cuda["rightlane_wpts"], mono_n),
            "stlp": mul_n(gt_stlp, mono_n)[:, None, :],
            "dense_valids": mul_n(torch.ones_like(batch_cuda["gt_high_level"]), mono_n),
            "gt_high_level": mul_n(batch_cuda["gt_high_level"], mono_n),
        }
    else:
        stl_input = {
            "neighbors": batch_cuda["neighbors_dense"],
            "currlane_wpts": batch_cuda["currlane_wpts_dense"],
            "leftlane_wpts": batch_cuda["leftlane_wpts_dense"],
            "rightlane_wpts": batch_cuda["rightlane_wpts_dense"],
            "stlp": batch_cuda["stlp_dense"],
            "dense_valids": batch_cuda["valids_dense"],
            "gt_high_level": batch_cuda["gt_high_level"],
        }
    if detach:
        stl_input = {k: v.detach() for k, v in stl_input.items()}
    if repeat_n is not None:
        stl_input = {k: v.repeat(repeat_n, *[1] * (v.dim() - 1)) for k, v in stl_input.items()}
    if dense_trajs is not None:
        stl_input["ego_traj"] = dense_trajs
    return stl_input


def augment_batch_data(batch, the_stlp, args, n_randoms=None, stlp_dense=None):
    """densify a scene batch to one row per chain (reference :724-754).  The replicated scene
    tensors are lazy; the kernels read the compact ``_pstl_pack`` instead."""
    if n_randoms is None:
        new_sample = False
        n_randoms = args.n_randoms
    else:
        new_sample = True
    m = n_randoms * 3
    bs = batch["currlane_wpts"].shape[0]
    if not isinstance(batch, LazyBatch):
        lb = LazyBatch(batch)
        batch = lb
    batch.set_lazy("neighbors_dense", lambda: dup(batch["neighbor_trajs_aug"], m))
    for k in ("curr", "left", "right"):
        batch.set_lazy("%slane_wpts_dense" % k, (lambda kk: (lambda: dup(batch["%slane_wpts" % kk], m)))(k))
    batch["stlp"] = the_stlp.unsqueeze(-2) if the_stlp is not None else None
    if stlp_dense is not None:
        batch["stlp_dense"] = stlp_dense
    elif args.load_stlp:
        if new_sample:
            batch["stlp_dense"] = batch["pre_stlp"].reshape(bs, args.n_randoms, 3, 6)[:, 0:1].repeat(
                1, args.sampling_size, 1, 1).reshape(bs * m, 1, 6)
        else:
            batch["stlp_dense"] = batch["pre_stlp"].reshape(bs * m, 1, 6)
    else:
        raise NotImplementedError("get_dense_stlp (pSTL calibration, reference :657-722) is out of scope: "
                                  "pass stlp_dense or use --load_stlp")
    valids = torch.cat([batch["curr_id"], batch["left_id"], batch["right_id"]], dim=-1)
    batch["valids_dense"] = dup(valids, n_randoms).reshape(bs * n_randoms, 3)
    batch["highlevel_dense"] = _mode_column(bs * n_randoms, valids.device)
    if "neighbor_trajs_aug" not in batch and "neighbors_traj" in batch:
        batch["neighbor_trajs_aug"] = batch["neighbors_traj"][..., :7]
    batch["_pstl_pack"] = ScenePack.from_batch(batch, batch["stlp_dense"], n_randoms)
    return batch


# ---------------------------------------------------------------------------------------
# scoring (reference :318-345) and best-of-K (:992-1013)
# ---------------------------------------------------------------------------------------

class _FusedScoreEgo(torch.autograd.Function):
    """scores of pre-rolled trajectories; differentiable w.r.t. ego_traj."""

    @staticmethod
    def forward(ctx, ego, progs, sv_parts, spec, mode, stlp):
        nei, lanes, rps = sv_parts
        N, T, W = ego.shape
        e = _nv.f32(ego)
        sv = _nv.make_scene_view(nei, lanes, rps)
        scores = torch.empty((N,), dtype=torch.float32, device=ego.device)
        L = _nv.lib()
        pa = _nv.prog_array(progs)
        ws = _nv.workspace(L.pstl_score_workspace_bytes(pa, N, T, 0), ego.device, "score")
        _nv.check(L.pstl_score_fused(pa, _nv.C.byref(sv), _nv.C.byref(spec), _nv.fptr(mode), None, None, 1,
                                     _nv.fptr(e), W, _nv.fptr(stlp), N, None, _nv.fptr(scores), None, None, None,
                                     _nv.ptr(ws), _nv.stream()), "pstl_score_fused")
        ctx.save_for_backward(e, mode, stlp)
        ctx.misc = (progs, sv_parts, spec, W)
        return scores

    @staticmethod
    def backward(ctx, g):
        e, mode, stlp = ctx.saved_tensors
        progs, (nei, lanes, rps), spec, W = ctx.misc
        N, T, _ = e.shape
        sv = _nv.make_scene_view(nei, lanes, rps)
        ge4 = torch.empty((N, T, 4), dtype=torch.float32, device=e.device)
        L = _nv.lib()
        pa = _nv.prog_array(progs)
        ws = _nv.workspace(L.pstl_score_workspace_bytes(pa, N, T, 1), e.device, "score")
        _nv.check(L.pstl_score_fused_bwd(pa, _nv.C.byref(sv), _nv.C.byref(spec), _nv.fptr(mode), None, None,
                                         _nv.fptr(e), W, _nv.fptr(stlp), N, _nv.fptr(_nv.f32(g)), None, None,
                                         _nv.fptr(ge4), _nv.ptr(ws), _nv.stream()), "pstl_score_fused_bwd")
        if W == 4:
            ge = ge4
        else:
            ge = torch.zeros((N, T, W), dtype=torch.float32, device=e.device)
            ge[..., :4] = ge4
        return ge, None, None, None, None, None


class _LazyScores(list):
    """scores_list of compute_stl_dense: the per-formula scores are evaluated on first access."""

    def __init__(self, fn):
        super().__init__()
        self._fn = fn

    def _fill(self):
        if self._fn is not None:
            super().extend(self._fn())
            self._fn = None

    def __getitem__(self, i):
        self._fill()
        return super().__getitem__(i)

    def __iter__(self):
        self._fill()
        return super().__iter__()

    def __len__(self):
        self._fill()
        return super().__len__()


def score_pack(pack, controls, args, progs, scaled=True, want=("best_score",)):
    """Fused rollout+predicates+STL(+best-of-K) on a ScenePack.
    controls: (C,N,T,2) or (N,T,2) physical controls (``scaled``) or raw mu/x (then scaled+clipped
    per normalize_diff).  Returns dict with the requested outputs among
    scores_all (C,N), best_score (N), best_idx (N), best_controls (N,T,2), traj (N,T+1,4)."""
    _check_fused_supported(args, "score_pack (the sampling pipeline)")
    c = controls if controls.dim() == 4 else controls.unsqueeze(0)
    c = _nv.f32(c)
    C_, N, T, _ = c.shape
    assert N == pack.N and T == pack.T
    dev = c.device
    out = {}
    if "scores_all" in want:
        out["scores_all"] = torch.empty((C_, N), dtype=torch.float32, device=dev)
    if "best_score" in want:
        out["best_score"] = torch.empty((N,), dtype=torch.float32, device=dev)
    if "best_idx" in want:
        out["best_idx"] = torch.empty((N,), dtype=torch.int32, device=dev)
    if "best_controls" in want:
        out["best_controls"] = torch.empty((N, T, 2), dtype=torch.float32, device=dev)
    if "traj" in want:
        out["traj"] = torch.empty((N, T + 1, 4), dtype=torch.float32, device=dev)
    spec = _spec(args) if scaled else _spec(args, args.mul_w_max, args.mul_a_max, int(bool(args.diffusion_clip)))
    sv = pack.view()
    L = _nv.lib()
    pa = _nv.prog_array(progs)
    ws = _nv.workspace(L.pstl_score_workspace_bytes(pa, N, T, 0), dev, "score")
    _nv.check(L.pstl_score_fused(pa, _nv.C.byref(sv), _nv.C.byref(spec), _nv.fptr(pack.mode), _nv.fptr(pack.state0),
                                 _nv.fptr(c), C_, None, 0, _nv.fptr(pack.stlp), N, _nv.fptr(out.get("scores_all")),
                                 _nv.fptr(out.get("best_score")), _nv.ptr(out.get("best_idx")),
                                 _nv.fptr(out.get("best_controls")), _nv.fptr(out.get("traj")), _nv.ptr(ws),
                                 _nv.stream()), "pstl_score_fused")
    return out


def compute_stl_dense(stl_input, stls_cac, stl_idx, mask, args, debug=False, tj_scores=None, scene=False):
    """evaluate the three formulas at t=0 and select by mode (reference :318-345).

    Returns (scores_list, scores, acc[, scene_acc | stl_input]) like upstream.  With the typed
    spec of ``build_stl_cache`` this is ONE fused kernel (predicates + formula, only the row's own
    formula — the reference evaluates all three and multiplies by one-hot masks, equal whenever the
    unused formulas are finite); custom AP lambdas, other anchor grids (--refined_nL / --refined_nW) and
    --collision_loss (whose extra signals the caller reads from stl_input) run predicates + generic interpreter."""
    ego = stl_input["ego_traj"]
    _nv.require_cuda(ego, "ego_traj")
    N, T = ego.shape[0], ego.shape[1]
    mode = _nv.f32(stl_idx[:, 0])
    fused_ok = _default_anchors(args) and getattr(args, "collision_loss", None) is None
    progs = _fused_programs(stls_cac, T) if fused_ok else None
    pack = stl_input.get("_pstl_pack") if isinstance(stl_input, dict) else None
    if progs is not None:
        if pack is not None:
            rep = stl_input.get("_pstl_repeat", 1)
            if rep == 1:
                parts, stlp = (pack.neighbors, pack.lanes, pack.rows_per_scene), pack.stlp
            else:  # candidate-major stacking: row r of candidate c -> scene of row r
                assert N == rep * pack.N
                parts, stlp = None, pack.stlp
        else:
            parts = (_nv.f32(stl_input["neighbors"]),
                     [_nv.f32(stl_input["%slane_wpts" % k]) for k in ("curr", "left", "right")], 1)
            stlp = _nv.f32(stl_input["stlp"].reshape(N, 6))
        if parts is None:
            chunks = [_FusedScoreEgo.apply(ego[c * pack.N:(c + 1) * pack.N], progs,
                                           (pack.neighbors, pack.lanes, pack.rows_per_scene), _spec(args),
                                           mode[c * pack.N:(c + 1) * pack.N].contiguous(), stlp) for c in range(rep)]
            scores = torch.cat(chunks, 0)
        else:
            scores = _FusedScoreEgo.apply(ego, progs, parts, _spec(args), mode, stlp)

        def all_formulas():
            outs = []
            for k in range(3):
                mk = torch.full_like(mode, float(k))
                if parts is None:
                    outs.append(torch.cat([_FusedScoreEgo.apply(ego[c * pack.N:(c + 1) * pack.N], progs,
                                                                (pack.neighbors, pack.lanes, pack.rows_per_scene),
                                                                _spec(args), mk[:pack.N].contiguous(), stlp)
                                           for c in range(rep)], 0))
                else:
                    outs.append(_FusedScoreEgo.apply(ego, progs, parts, _spec(args), mk, stlp))
            outs.append(outs[-1].detach() * 0.0 + 1.0)
            return outs

        scores_list = _LazyScores(all_formulas)
    else:
        stl_input = prep_stl_cache(stl_input, args)
        res = [f(stl_input, args.smoothing_factor) for f in stls_cac]
        scores_list = [r[:, 0] for r in res]
        scores_list.append(scores_list[-1].detach() * 0.0 + 1.0)
        scores = get_stl_scores(scores_list, stl_idx[:, 0])
    mask_flat = mask.reshape(-1)
    if getattr(args, "oracle_filter", False) and tj_scores is not None:
        cube = torch.max(tj_scores.reshape(-1, args.n_randoms, 3), dim=1, keepdim=True)[0]
        tj_val = ((cube > 0).float()).repeat(1, args.n_randoms, 1).reshape(-1)
        acc = mask_mean((scores > 0).float(), mask_flat * tj_val)
    else:
        acc = mask_mean((scores > 0).float(), mask_flat)
    if debug:
        if progs is not None and "x2curr_d" not in stl_input:
            stl_input = prep_stl_cache(stl_input, args)
        return scores_list, scores, acc, stl_input
    if scene:
        scores_cube = scores.reshape(-1, args.n_randoms, 3)
        mask_cube = mask.reshape(-1, args.n_randoms, 3)
        scene_acc = mask_mean((torch.max(scores_cube, dim=1)[0] > 0).float(), mask_cube[:, 0, :])
        return scores_list, scores, acc, scene_acc
    return scores_list, scores, acc


# ---------------------------------------------------------------------------------------
# training losses of the RefineNet step (reference :370-478, diffusion + rect_head branch)
# ---------------------------------------------------------------------------------------

class _RefineLosses(torch.autograd.Function):
    """(loss, terms) = pstl_refine_losses(rect_controls, scores); the kernel returns d loss / d rect_controls and
    d loss / d scores with the value, backward scales them by the incoming gradient of ``loss``.  ``terms``
    (loss_stl, loss_reg, loss_diversity, extra_loss_reg, ...) are reported values, not differentiable."""

    @staticmethod
    def forward(ctx, rect, scores, nn_controls, valid, cfg):
        _nv.require_cuda(rect, "rect_controls")
        r, n = _nv.f32(rect.reshape(rect.shape[0], -1)), _nv.f32(nn_controls.reshape(rect.shape[0], -1))
        sc, vl = _nv.f32(scores.reshape(-1)), _nv.f32(valid.reshape(-1))
        assert r.shape[0] == cfg.n_scenes * cfg.S * 3 == sc.shape[0] == vl.shape[0] and r.shape[1] == 2 * cfg.nt
        L = _nv.lib()
        losses = torch.empty((8,), dtype=torch.float32, device=rect.device)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        d_rect = torch.empty_like(r) if need else None
        d_sc = torch.empty_like(sc) if need else None
        ws = _nv.workspace(L.pstl_refine_losses_workspace_bytes(_nv.C.byref(cfg)), rect.device, "losses")
        _nv.check(L.pstl_refine_losses(_nv.C.byref(cfg), _nv.fptr(r), _nv.fptr(n), _nv.fptr(sc), _nv.fptr(vl),
                                       _nv.fptr(losses), _nv.fptr(d_rect), _nv.fptr(d_sc), _nv.ptr(ws), _nv.stream()),
                  "pstl_refine_losses")
        ctx.grads = (d_rect, d_sc, rect.shape, scores.shape)
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g, _g_terms):
        d_rect, d_sc, rs, ss = ctx.grads
        return ((g * d_rect).reshape(rs) if ctx.needs_input_grad[0] else None,
                (g * d_sc).reshape(ss) if ctx.needs_input_grad[1] else None, None, None, None)


def loss_cfg(args, bs, S=None):
    """pstl_loss_cfg from the parsed flags."""
    return _nv.LossCfg(n_scenes=bs, S=S or args.n_randoms, nt=args.nt, n_shards=args.n_shards,
                       diverse_loss=int(bool(args.diverse_loss)), diverse_detach=int(bool(args.diverse_detach)),
                       w_max=args.mul_w_max, a_max=args.mul_a_max, stl_nn_thres=args.stl_nn_thres,
                       stl_weight=args.stl_weight, diversity_scale=args.diversity_scale,
                       diversity_weight=args.diversity_weight, rect_reg_loss=args.rect_reg_loss,
                       extra_rect_reg=(args.extra_rect_reg or 0.0))


def evaluate_all_scores(scores, gt_labels, valid_mask, n_randoms):
    """reference :347-368: the per-scene score vectors grouped by in-label / out-of-label lane."""
    names = ("curr", "left", "right")
    keys = ["in_label_scores", "out_label_scores"] + ["%s_label_%s_scores" % (io, n) for io in ("in", "out") for n in names]
    res = {k: [] for k in keys}
    bs = gt_labels.shape[0]
    # one device->host copy of the small label / validity tables and one unbind for the (scene, mode) score vectors
    rows = scores.detach().reshape(bs, n_randoms, 3).permute(0, 2, 1).reshape(bs * 3, n_randoms).unbind(0)
    vm = valid_mask.reshape(bs, n_randoms, 3)[:, 0].tolist()
    lab = gt_labels.reshape(bs, -1)[:, 0].tolist()
    for i in range(bs):
        if lab[i] < 3:
            for j in range(3):
                if vm[i][j] > 0:
                    io = "in" if lab[i] == j else "out"
                    res["%s_label_scores" % io].append(rows[i * 3 + j])
                    res["%s_label_%s_scores" % (io, names[j])].append(rows[i * 3 + j])
    return res


def compute_policy_loss(batch_cuda, nn_stlp, stls_cac, nn_trajs, rect_trajs, dense_trajs, args, diffusion_extras=None,
                        vae_extras=None, dbgs_extras=None, bc_extras=None, nn_controls_adj=None,
                        nn_controls_list_adj=None, opt_controls=None):
    """The loss of one --diffusion training step (reference :370-478; the vae / bc / single-sample branches are not
    built).  Scores of the trajectories that are trained (rect_trajs with --rect_head) come from the fused scoring
    kernel; with --rect_head every loss term and its gradient w.r.t. rect_controls and the scores come from ONE native
    call (pstl_refine_losses) instead of the ~40 autograd nodes upstream records.  Returns (rd, all_scores)."""
    if diffusion_extras is None or vae_extras is not None or bc_extras is not None:
        raise NotImplementedError("compute_policy_loss: only the diffusion branch is built")
    bs = batch_cuda["ego_traj"].shape[0]
    S = args.n_randoms
    self_trajs = rect_trajs if args.rect_head else nn_trajs
    noised_a, est_cmds_a, highlevel_dense, dense_scores, dense_valids, epi, raw_noise, nn_controls, steps, rect_controls = \
        diffusion_extras
    rd = {"avg_speed": torch.mean(self_trajs[..., :-1, 3]), "avg_speed_gt": torch.mean(batch_cuda["ego_traj"][..., 3])}
    stl_input = pre_prepare_stl_cache(batch_cuda, dense_trajs=self_trajs[:, :-1])
    valid_mask = batch_cuda["valids_dense"].reshape(-1)
    _, scores, acc = compute_stl_dense(stl_input, stls_cac, batch_cuda["highlevel_dense"], valid_mask, args)
    stl_input_gt = {"ego_traj": batch_cuda["ego_traj"], "neighbors": batch_cuda["neighbor_trajs_aug"],
                    "stlp": batch_cuda["stlp"]}
    for k in ("currlane_wpts", "leftlane_wpts", "rightlane_wpts"):
        stl_input_gt[k] = batch_cuda[k]
    gt_hl = batch_cuda["gt_high_level"]
    _, scores_gt, acc_gt = compute_stl_dense(stl_input_gt, stls_cac, gt_hl, (gt_hl[:, 0] != 3).float(), args)
    all_scores = evaluate_all_scores(scores, gt_hl, valid_mask, S)
    rd.update(acc=acc, acc_gt=acc_gt, scores=scores, scores_all=scores, scores_gt=scores_gt, scores_gt_all=scores_gt)
    if getattr(args, "stl_bc_mask", False):
        m = (dense_scores * dense_valids > 0).float().reshape(bs * S * 3, 1)
        rd["loss_diffusion"] = mask_mean(torch.square(raw_noise - est_cmds_a), m)
    else:
        rd["loss_diffusion"] = torch.mean(torch.square(raw_noise - est_cmds_a))
    loss_coll = None
    if getattr(args, "collision_loss", None) is not None:
        # TrafficSim-style collision term (reference :416-420) on the extra signals prep_stl_cache added to stl_input
        coll_dist = torch.relu(1 - stl_input["min_centroid_d"] / torch.clip(stl_input["radius_sum"], 1e-1))
        loss_coll = torch.mean(torch.clip(torch.sum(coll_dist, dim=-1), max=1)) * args.collision_loss
    if args.rect_head:
        loss, terms = _RefineLosses.apply(rect_controls, scores, nn_controls.detach(), valid_mask, loss_cfg(args, bs, S))
        rd["loss"], rd["loss_stl"], rd["loss_reg"] = loss, terms[1], terms[2]
        rd["loss_coll"] = terms[1] * 0 if loss_coll is None else loss_coll
        if args.diverse_loss:
            rd["loss_diversity"] = terms[3]  # upstream's --diverse_loss total leaves loss_coll out (:466)
        else:
            rd["extra_loss_reg"] = terms[4]
            if loss_coll is not None:
                rd["loss"] = rd["loss"] + loss_coll
    else:
        rd["loss_stl"] = mask_mean(torch.relu(args.stl_nn_thres - scores), valid_mask) * args.stl_weight
        rd["loss_coll"] = rd["loss_stl"] * 0 if loss_coll is None else loss_coll
        rd["loss"] = rd["loss_stl"] + rd["loss_diffusion"] + rd["loss_coll"]
    return rd, all_scores


# ---------------------------------------------------------------------------------------
# sampler (reference :528-655)
# ---------------------------------------------------------------------------------------

def get_diffusion_coeffs(args):
    """cosine / linear schedule (reference :528-537): (beta, alpha, alpha_hat) on the GPU."""
    if args.cos:
        t = torch.linspace(0, 1, args.diffusion_steps + 1)
        alpha_bar = torch.cos((t + 0.008) / 1.008 * np.pi / 2) ** 2
        beta = torch.clip(1 - alpha_bar[1:] / alpha_bar[:-1], 0, 0.999) * 0.2
    else:
        beta = torch.linspace(args.beta_start, args.beta_end, args.diffusion_steps)
    alpha = 1.0 - beta
    alpha_hat = torch.cumprod(alpha, dim=0)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    return (beta.to(dev), alpha.to(dev), alpha_hat.to(dev))


def normalize_diff(x, n, nt, w_max, a_max, clip):
    """reference :647-655 (elementwise; the sampler fuses it, this is the standalone form)."""
    x = x.reshape(n, nt, 2)
    w, a = x[..., 0] * w_max, x[..., 1] * a_max
    if clip:
        w, a = torch.clip(w, -w_max, w_max), torch.clip(a, -a_max, a_max)
    return torch.stack([w, a], dim=-1)


class IterateList:
    """final_list of diffusion_rollout (reference :633-634): ``steps`` entries x_T..x_0, of which only
    the last ``K`` are stored (the reference keeps all 100 = 3.1 GB at 196,608 chains)."""

    def __init__(self, steps, kept, first=None):
        self.steps, self.kept = steps, kept  # kept (K,N,T,2), chronological
        self.first = first                   # normalised x_T (entry 0), when the caller kept it

    def __len__(self):
        return self.steps

    def _one(self, i):
        if i < 0:
            i += self.steps
        if i == 0 and self.first is not None:
            return self.first
        j = i - (self.steps - self.kept.shape[0])
        if j < 0 or i >= self.steps:
            raise IndexError("iterate %d was not kept (only the last %d are; pass keep_all_iterates)"
                             % (i, self.kept.shape[0]))
        return self.kept[j]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(j) for j in range(*i.indices(self.steps))]
        return self._one(i)

    def stacked_last(self, k):
        return self.kept[self.kept.shape[0] - k:]


_call_counter = [0]
_sched_cache = {}
KERNEL_TIMER = None  # bench hook: callable(name, is_start) recording CUDA events around the native sampler call


def _host_schedule(beta, alpha, alpha_hat):
    """(3,steps) fp32 host copy of the schedule, read back from the device once per coeffs tuple
    (a per-call .cpu() would be a device synchronisation in the middle of the pipeline)."""
    key = (beta.data_ptr(), alpha.data_ptr(), alpha_hat.data_ptr(), beta._version, beta.numel())
    s = _sched_cache.get(key)
    if s is None:
        s = torch.stack([beta, alpha, alpha_hat], 0).to("cpu", torch.float32).contiguous()
        _sched_cache.clear()
        _sched_cache[key] = s
    return s


def guidance_step(pack, mu, stls_cac, args, beta_t, it=0, state=None, maximize=False, n_total=None):
    """ONE iteration of the guidance block of a reverse step on ``mu`` (N, nt, 2), in place (reference :605-626, the
    body of ``for j in range(guidance_niters)``): rollout of ``mu * (w_max, a_max)`` -> STL scores -> loss
    ``mask_mean(relu(thres - score), valid)`` -> Adam step (+ the |delta| clip to ``beta_t`` from the second iteration,
    see the aliasing note in DESIGN.md).  ``state`` = (m, v, anchor) Adam moments / anchor, created zeroed when None
    (pass it back for ``it`` >= 1).  Returns (mu, grad (N, nt, 2) of the loss w.r.t. mu, scores are not returned).
    The sampler runs exactly this call inside ``pstl_denoiser_sample``; it is exposed for tests and custom loops."""
    _check_fused_supported(args, "guidance_step")
    N, T = pack.N, pack.T
    mu = mu.reshape(N, T, 2)
    _nv.require_cuda(mu, "mu")
    if mu.dtype != torch.float32 or not mu.is_contiguous():
        raise ValueError("mu must be contiguous float32 (it is updated in place)")
    progs = _fused_programs(stls_cac, T)
    if progs is None:
        raise NotImplementedError("guidance needs the typed spec of build_stl_cache")
    L = _nv.lib()
    pa = _nv.prog_array(progs)
    sv, sp = pack.view(), _spec(args, args.mul_w_max, args.mul_a_max, 0)
    if state is None:
        state = tuple(torch.zeros_like(mu) for _ in range(3))
    ws = _nv.workspace(N * T * 2 * 4 + L.pstl_score_workspace_bytes(pa, N, T, 1), mu.device, "guidance")
    n_total = float(N if n_total is None else n_total)
    inv_norm = 1.0 / (n_total * max(float(pack.valid.mean().item()), 1e-2))
    thres = 100.0 if maximize else args.stl_nn_thres
    _nv.check(L.pstl_guidance_step(pa, _nv.C.byref(sv), _nv.C.byref(sp), _nv.fptr(pack.mode), _nv.fptr(pack.state0),
                                   _nv.fptr(pack.stlp), _nv.fptr(pack.valid), N, _nv.C.c_float(thres),
                                   _nv.C.c_float(inv_norm), None, _nv.C.c_float(args.guidance_lr),
                                   _nv.C.c_float(beta_t), int(it), _nv.fptr(mu), _nv.fptr(state[0]), _nv.fptr(state[1]),
                                   _nv.fptr(state[2]), _nv.ptr(ws), _nv.stream()), "pstl_guidance_step")
    grad = ws[:N * T * 2 * 4].view(torch.float32).reshape(N, T, 2).clone()
    return mu, grad, state


def guidance_step_mask(args):
    """which reverse steps i (index into a ``diffusion_steps``-long uint8 array) run STL guidance — the trigger chain
    of reference :589-598: ``i_val in guidance_sets``, else ``i_val % guidance_freq == 0``, else ``i <= guidance_before``
    with ``i_val = diffusion_steps-1-i`` under --guidance_reverse (the last rule reads ``i`` itself, as upstream)."""
    steps = args.diffusion_steps
    mask = np.zeros(steps, dtype=np.uint8)
    for i in range(1, steps):
        i_val = steps - 1 - i if args.guidance_reverse else i
        if args.guidance_sets is not None:
            hit = i_val in args.guidance_sets
        elif args.guidance_freq is not None:
            hit = i_val % args.guidance_freq == 0
        else:
            hit = i <= args.guidance_before
        mask[i] = 1 if hit else 0
    return mask


def diffusion_rollout(noise, net, batch_cuda, highlevel_dense, feature, args, coeffs=None, fastforward=False,
                      n_randoms=None, return_feature=False, mono=False, tmp_stlp=None, guidance_extras=None,
                      maximize=False, scene_feature_only=False):
    """DDPM reverse loop i = steps-1..1 with t == i (reference :557-645): ONE native call that runs all
    steps (eps-MLP, posterior update, noise, optional STL guidance) instead of ~100 launches per step.

    ``noise`` gives only shape/device, as upstream (:563).  Deterministic mode: ``args.inject_noise`` =
    [x_T, z_1, ...] tensors (N,2nt) consumed in upstream's randn_like order; otherwise x_T and z are drawn by the
    kernels' Philox stream (``args.seed``)."""
    _nv.require_cuda(noise, "noise")
    n = noise.shape[0]
    nt, T2, steps = args.nt, args.nt * 2, args.diffusion_steps
    beta, alpha, alpha_hat = coeffs
    net.eval()
    bs = batch_cuda["ego_traj"].shape[0]
    if n_randoms is None:
        n_randoms = args.n_randoms
    if fastforward:
        # reference :567: the reverse loop is skipped, the "sample" is the initial noise x_T (a training-time switch
        # that keeps the call's return shape while saving the sampling cost on epochs that are not visualised)
        inj = getattr(args, "inject_noise", None)
        x_T = _nv.f32(inj[0]) if inj is not None else torch.randn((n, T2), device=noise.device)
        final = normalize_diff(x_T, n, nt, args.mul_w_max, args.mul_a_max, args.diffusion_clip)
        dense_feature = feature
        if args.diff_full:
            fl = IterateList(1, final.unsqueeze(0))
            return (final, dense_feature, fl) if return_feature else (final, fl)
        return (final, dense_feature) if return_feature else final
    scene_feat = getattr(feature, "_pstl_scene_feat", None) if feature is not None else None
    if scene_feat is None:
        if feature is not None:
            scene_feat = feature.reshape(bs, -1, feature.shape[-1])[:, 0].contiguous()
        else:
            with torch.no_grad():
                scene_feat = net.encode_feat(batch_cuda)
    rows_per_scene = n // bs
    if mono:
        # --gt_data_training layout (reference :570-572, nusc_model.py:124-128): one row per (scene, sample); the scene's
        # high-level mode (bs,1) and ground-truth pSTL parameters tmp_stlp (bs,6) are shared by its n // bs rows
        if feature is None or tmp_stlp is None:
            raise ValueError("mono sampling takes prev_feature (bs, k) and tmp_stlp (bs, 6), as upstream")
        hl = _nv.f32(highlevel_dense.reshape(bs, 1).expand(bs, rows_per_scene).reshape(n))
        stlp = _nv.f32(tmp_stlp.reshape(bs, 1, 6).expand(bs, rows_per_scene, 6).reshape(n, 6))
        scene_feat = _nv.f32(feature.reshape(bs, -1))
    else:
        hl = _nv.f32(highlevel_dense.reshape(n))
        stlp = _nv.f32(batch_cuda["stlp_dense"].reshape(n, 6))
    inj = getattr(args, "inject_noise", None)
    if inj is not None:
        x_T = _nv.f32(inj[0])
        z = torch.stack([_nv.f32(t) for t in inj[1:steps - 1]], 0).contiguous() if steps > 2 else None
    else:
        # x_T ~ N(0,1): drawn by the sampler from its own Philox stream (no separate randn launch / HBM round trip)
        x_T = None
        z = None
    keep_all = bool(getattr(args, "refinement", False) or getattr(args, "keep_all_iterates", False))
    K = steps - 1 if keep_all else max(1, int(args.multi_cands or 1))
    K = min(K, steps - 1)
    if keep_all and x_T is None:
        x_T = torch.randn((n, T2), device=noise.device)  # entry 0 of final_list (x_T itself) must be readable
    iterates = torch.empty((K, n, nt, 2), dtype=torch.float32, device=noise.device)
    handle = net.native_handle(getattr(args, "precision", "fp32"))
    L = _nv.lib()
    gcfg = None
    keepalive = []
    if args.guidance:
        _check_fused_supported(args, "--guidance")
        new_batch, states_flat_new, stls_cac = guidance_extras
        pack = new_batch.get("_pstl_pack")
        if pack is None:
            pack = ScenePack.from_batch(new_batch, batch_cuda["stlp_dense"], n_randoms)
        progs = _fused_programs(stls_cac, nt)
        if progs is None:
            raise NotImplementedError("guidance needs the typed spec of build_stl_cache")
        valid = pack.valid
        sv = pack.view()
        sp = _spec(args, args.mul_w_max, args.mul_a_max, 0)
        pa = _nv.prog_array(progs)
        n_total = float(getattr(args, "guidance_n_total", n))
        gcfg = _nv.GuidanceCfg()
        mv = getattr(args, "guidance_mean_valid", None)
        if mv is None:
            # the loss normaliser 1 / (N clip(mean(valid), 1e-2)) (reference :23-27, 616-619) stays on the device: no host
            # read-back in the middle of the batch, and a captured graph follows each batch's own validity mean
            inv_dev = (1.0 / (n_total * torch.clip(valid.mean().double(), 1e-2))).float().reshape(1)
            gcfg.inv_norm_dev = inv_dev.data_ptr()
            gcfg.inv_norm = 0.0
            keepalive.append(inv_dev)
        else:
            gcfg.inv_norm = 1.0 / (n_total * max(float(mv), 1e-2))
        s0 = _nv.f32(states_flat_new)
        keepalive.append(s0)
        gcfg.valid, gcfg.state0 = valid.data_ptr(), s0.data_ptr()
        gcfg.progs = _nv.C.cast(pa, _nv.C.POINTER(_nv.C.c_void_p))
        gcfg.scenes, gcfg.sp = _nv.C.pointer(sv), _nv.C.pointer(sp)
        gcfg.before, gcfg.niters = int(min(args.guidance_before, steps - 1)), int(args.guidance_niters)
        step_mask = guidance_step_mask(args)
        gcfg.step_mask = step_mask.ctypes.data
        keepalive.append(step_mask)
        gcfg.lr = args.guidance_lr
        gcfg.thres = 100.0 if maximize else args.stl_nn_thres
        keepalive += [pack, sv, sp, pa, progs, states_flat_new]
    ws_bytes = L.pstl_denoiser_workspace_bytes(handle, n, bs, _nv.C.byref(gcfg) if gcfg is not None else None)
    ws = _nv.workspace(ws_bytes, noise.device, "denoiser")
    sched = _host_schedule(beta, alpha, alpha_hat)
    temb = net.time_table(steps, noise.device)
    _call_counter[0] += 1
    if KERNEL_TIMER is not None:
        KERNEL_TIMER("sampler", True)
    _nv.check(L.pstl_denoiser_sample(
        handle, _nv.fptr(_nv.f32(scene_feat)), bs, rows_per_scene, _nv.fptr(hl), _nv.fptr(stlp), n,
        _nv.C.c_void_p(sched.data_ptr()), _nv.fptr(temb), steps, _nv.fptr(x_T), _nv.fptr(z),
        _nv.C.c_uint64(int(getattr(args, "seed", 0)) & (2 ** 64 - 1)), _nv.C.c_uint64(_call_counter[0] * 1000),
        _nv.C.c_float(args.mul_w_max), _nv.C.c_float(args.mul_a_max), int(bool(args.diffusion_clip)), K,
        _nv.C.byref(gcfg) if gcfg is not None else None, _nv.fptr(iterates), None, _nv.ptr(ws), _nv.stream()),
        "pstl_denoiser_sample")
    if KERNEL_TIMER is not None:
        KERNEL_TIMER("sampler", False)
    diffused_result = iterates[-1]
    dense_feature = None
    if return_feature:
        k = scene_feat.shape[-1]
        if scene_feature_only:
            # callers inside this package only read the per-scene rows: skip the (N, 224) replication
            # (176 MB at 196,608 chains) that upstream's dense feature costs
            dense_feature = scene_feat.reshape(bs, 1, k).expand(bs, rows_per_scene, k)
        else:
            dense_feature = scene_feat.reshape(bs, 1, k).expand(bs, rows_per_scene, k).reshape(-1, k)
        dense_feature._pstl_scene_feat = scene_feat
    if args.diff_full:
        first = normalize_diff(x_T, n, nt, args.mul_w_max, args.mul_a_max, args.diffusion_clip) if keep_all else None
        final_list = IterateList(steps, iterates, first)
        return (diffused_result, dense_feature, final_list) if return_feature else (diffused_result, final_list)
    return (diffused_result, dense_feature) if return_feature else diffused_result


# ---------------------------------------------------------------------------------------
# open-loop sampling test (reference :890-1183; the timed region :957-1105)
# ---------------------------------------------------------------------------------------

def accuracy(scores, valid, bs, S):
    """(acc, scene_acc) of the report line (reference :336-343): mask_mean of (score > 0) over the chains and of
    (best sample of a (scene, mode) > 0) over the scenes' lanes — two small launches instead of a dozen tensor ops."""
    _nv.require_cuda(scores, "scores")
    part = torch.empty((bs, 4), dtype=torch.float32, device=scores.device)
    out = torch.empty((2,), dtype=torch.float32, device=scores.device)
    _nv.check(_nv.lib().pstl_accuracy(_nv.fptr(_nv.f32(scores.reshape(-1))), _nv.fptr(_nv.f32(valid.reshape(-1))), bs, S,
                                      _nv.fptr(part), _nv.fptr(out), _nv.stream()), "pstl_accuracy")
    return out[0], out[1]


@torch.no_grad()
def sample_and_score(net, batch_cuda, stls_cac, coeffs, args):
    """The timed region of run_sampling_test for the diffusion + RefineNet path:
    augment -> sampler -> best-of-K -> RefineNet (+n_rolls) -> final rollout + scores."""
    S = args.sampling_size
    bs = batch_cuda["ego_traj"].shape[0]
    N = bs * S * 3
    new_batch = LazyBatch({k: batch_cuda[k] for k in ("ego_traj", "neighbors", "currlane_wpts", "leftlane_wpts",
                                                       "rightlane_wpts", "curr_id", "left_id", "right_id",
                                                       "gt_high_level", "pre_stlp") if k in batch_cuda})
    new_batch["neighbor_trajs_aug"] = batch_cuda["neighbors_traj"][..., :7]
    new_batch = augment_batch_data(new_batch, None, args, n_randoms=S)
    pack = new_batch["_pstl_pack"]
    highlevel_new = new_batch["highlevel_dense"]
    noise = torch.empty((N, args.nt * 2), device=highlevel_new.device)
    guidance_extras = (new_batch, pack.state0, stls_cac) if args.guidance else None
    progs = _fused_programs(stls_cac, args.nt)
    res = diffusion_rollout(noise, net, new_batch, highlevel_new, None, args, coeffs, n_randoms=S,
                            return_feature=True, guidance_extras=guidance_extras, scene_feature_only=True)
    if args.diff_full:
        nn_controls, feature, nn_list = res
    else:
        nn_controls, feature = res
        nn_list = None
    out = {"final_iterate": nn_controls, "iterates": nn_list}
    if args.rect_head and not args.not_use_rect:
        if args.multi_cands is not None:
            cand = nn_list.stacked_last(args.multi_cands)  # (K,N,T,2) physical controls
            r = score_pack(pack, cand, args, progs, want=("scores_all", "best_score", "best_idx", "best_controls"))
            nn_controls, prev_scores = r["best_controls"], r["best_score"]
            out.update(cand_scores=r["scores_all"], best_idx=r["best_idx"], best_controls=nn_controls)
        else:
            prev_scores = score_pack(pack, nn_controls, args, progs)["best_score"]
        if not (args.multi_cands is not None and args.no_refinenet):
            nn_controls = net.rect_forward(feature, highlevel_new, new_batch["stlp_dense"][:, 0], nn_controls, prev_scores)
        if args.n_rolls is not None:
            for _ in range(args.n_rolls):
                sc = score_pack(pack, nn_controls, args, progs)["best_score"]
                nn_controls = net.rect_forward(feature, highlevel_new, new_batch["stlp_dense"][:, 0], nn_controls, sc)
        if getattr(args, "refinement", False):
            nn_controls = refine_by_mixing(pack, nn_controls, nn_list, stls_cac, args)
            out["refinement"] = True
    r = score_pack(pack, nn_controls, args, progs, want=("best_score", "traj"))
    scores, nn_trajs = r["best_score"], r["traj"]
    acc, scene_acc = accuracy(scores, pack.valid, bs, S)
    # the plan a caller would execute per scene: the best-scoring chain over its samples and valid lane modes
    pick = torch.where(pack.valid > 0, scores, torch.full_like(scores, -1e4)).reshape(bs, S * 3).argmax(dim=1)
    rows = pick + torch.arange(bs, device=pick.device) * (S * 3)
    out.update(controls=nn_controls, scores=scores, trajs=nn_trajs, acc=acc, scene_acc=scene_acc, pack=pack,
               scene_pick=rows, scene_plan=nn_controls.reshape(N, args.nt, 2)[rows])
    return out


REFINEMENT_ITERATES = {2: [0], 3: [80, 95], 4: [80, 90, 95], 6: [0, 50, 80, 90, 95], 8: [0, 50, 80, 85, 90, 95, 98],
                       10: [0, 50, 80, 85, 90, 95, 96, 97, 98],
                       20: [0, 10, 30, 50, 60, 70, 75, 80, 85, 90, 91, 92, 93, 94, 95, 96, 97, 98, 99]}


def refine_by_mixing(pack, nn_controls, nn_controls_list, stls_cac, args, K=8, n_iters=50, stl_thres=0.0005, lr=3e-1):
    """--refinement (reference nusc_train.py:1034-1071): rows whose final controls still violate the spec are replaced
    by a convex mix of those controls and K-1 earlier iterates of their own chain (REFINEMENT_ITERATES[K] indexes
    final_list); the softmax logits of the mix take ``n_iters`` Adam steps on mask_mean(relu(stl_thres - score), valid).
    Upstream's constants (K=8, 50 iterations, lr 0.3) are the defaults.  Every iteration is the fused reverse-mode
    scorer (score + d score / d controls in one launch) plus a handful of tensor ops on the (N, K) logits."""
    mix = MixingProblem(pack, nn_controls, nn_controls_list, stls_cac, args, K, stl_thres)
    optim = None
    with torch.enable_grad():
        lamdas = torch.ones(pack.N, K, device=nn_controls.device, requires_grad=True)
        optimizer = torch.optim.Adam([lamdas], lr=lr)
        for _ in range(n_iters):
            optim, _ = mix.backward_into(lamdas)
            optimizer.step()
    return optim.detach().reshape(pack.N, pack.T, 2)


class MixingProblem:
    """The objective of --refinement for one batch: ``backward_into(lamdas)`` evaluates the mixed controls of the current
    logits, their scores and the gradient of mask_mean(relu(stl_thres - score), valid) w.r.t. the logits (left in
    ``lamdas.grad``) — what one iteration of the reference's loop computes before ``optimizer.step()``."""

    def __init__(self, pack, nn_controls, nn_controls_list, stls_cac, args, K=8, stl_thres=0.0005):
        N, T = pack.N, pack.T
        self.pack, self.args, self.stl_thres = pack, args, stl_thres
        self.progs = _fused_programs(stls_cac, T)
        base = torch.stack([nn_controls.detach()] + [nn_controls_list[i].detach() for i in REFINEMENT_ITERATES[K]], dim=1)
        self.base = base.reshape(N, K, T * 2)                                 # (N, K, 2T) physical controls
        self.scores0 = score_pack(pack, nn_controls, args, self.progs)["best_score"]
        self.violated = ((self.scores0 <= 0) & (pack.valid > 0)).float().reshape(N, 1)
        self.keep = nn_controls.detach().reshape(N, T * 2) * (1 - self.violated)
        L = _nv.lib()
        self.pa = _nv.prog_array(self.progs)
        self.sv, self.sp = pack.view(), _spec(args)
        self.ws = _nv.workspace(L.pstl_score_workspace_bytes(self.pa, N, T, 1), base.device, "score")
        # d loss / d score = -valid / (N clip(mean(valid), 1e-2)) on the rows where stl_thres - score > 0 (relu'), else 0;
        # the reverse-mode scorer is linear in it, so it runs with the first factor and its rows are masked afterwards
        self.g_s = (-pack.valid / (N * torch.clip(pack.valid.mean(), 1e-2))).contiguous()
        self.scores = torch.empty((N,), dtype=torch.float32, device=base.device)
        self.g_u = torch.empty((N, T, 2), dtype=torch.float32, device=base.device)

    def backward_into(self, lamdas):
        pack, N, T = self.pack, self.pack.N, self.pack.T
        ratios = torch.softmax(lamdas, dim=-1)
        optim = self.keep + self.violated * torch.einsum("nk,nkd->nd", ratios, self.base)
        oc = optim.detach().reshape(N, T, 2).contiguous()
        _nv.check(_nv.lib().pstl_score_fused_bwd(self.pa, _nv.C.byref(self.sv), _nv.C.byref(self.sp), _nv.fptr(pack.mode),
                                                 _nv.fptr(pack.state0), _nv.fptr(oc), None, 0, _nv.fptr(pack.stlp), N,
                                                 _nv.fptr(self.g_s), _nv.fptr(self.scores), _nv.fptr(self.g_u), None,
                                                 _nv.ptr(self.ws), _nv.stream()), "pstl_score_fused_bwd")
        active = (self.stl_thres - self.scores > 0).float().reshape(N, 1)
        if lamdas.grad is not None:
            lamdas.grad = None
        optim.backward(self.g_u.reshape(N, T * 2) * active)
        return optim, self.scores


def diffusion_prep(dense_controls, n_randoms, coeffs=None, args=None, mono=False):
    """reference :539-555: normalised commands, one timestep per row in [1, diffusion_steps), the forward-noised
    commands.  Returns (noise, t (n,1), None, noised) like upstream."""
    if mono:
        raise NotImplementedError("mono (ground-truth-only) training is not built")
    n = dense_controls.shape[0] * n_randoms * 3
    cmd = dense_controls.reshape(n, args.nt, 2) / torch.tensor([args.mul_w_max, args.mul_a_max], device=dense_controls.device)
    cmd = cmd.reshape(n, args.nt * 2)
    noise = torch.normal(0, 1, (n, args.nt * 2), device=cmd.device)
    beta, alpha, alpha_hat = coeffs
    t = torch.randint(low=1, high=args.diffusion_steps, size=(n,), device=cmd.device)
    ah = alpha_hat.to(cmd.device)[t]
    return noise, t[:, None], None, torch.sqrt(ah)[:, None] * cmd + torch.sqrt(1 - ah)[:, None] * noise


def train_step_ddpm(net, batch_cuda, coeffs, args, optimizer=None, gt_stlp=None, prep=None):
    """One training iteration of the denoiser stage (README step 1; reference nusc_train.py:1352-1356, :436, :1523-1525):
    ``diffusion_prep`` on the stored (traj-opt) controls, ``net(...)`` with one timestep per row, the eps-prediction
    loss, backward (policy_net on the native kernels, the scene encoders through autograd from the per-scene feature
    gradient) and the optimiser step (``Adam(net.parameters())`` upstream).  ``prep`` = (noise, t, noised) overrides the
    draw.  The loss is upstream's: ``mask_mean(square(noise - est), tj_scores_prior * valids_dense > 0)`` under
    ``stl_bc_mask`` (forced by the parser), the plain mean otherwise.  With ``--stl_weight 0`` (the README command) this
    is the whole loss; the STL term of a sampled rollout that upstream adds for a non-zero weight is not built.
    Returns ``rd``."""
    if float(args.stl_weight) != 0.0:
        raise NotImplementedError("denoiser stage: only --stl_weight 0.0 (README step 1) is built")
    S = args.n_randoms
    bs = batch_cuda["ego_traj"].shape[0]
    nb = LazyBatch({k: batch_cuda[k] for k in ("ego_traj", "neighbors", "currlane_wpts", "leftlane_wpts", "rightlane_wpts",
                                               "curr_id", "left_id", "right_id", "gt_high_level", "pre_stlp", "params",
                                               "tj_scores_prior")
                    if k in batch_cuda})
    nb["neighbor_trajs_aug"] = batch_cuda["neighbors_traj"][..., :7]
    if gt_stlp is None:
        gt_stlp = batch_cuda["pre_stlp"].reshape(bs, -1, 3, 6)[:, 0, 0]
    nb = augment_batch_data(nb, gt_stlp, args)
    if prep is None:
        noise, steps, _, noised = diffusion_prep(nb["params"], S, coeffs, args)
    else:
        noise, steps, noised = prep
    est, feature = net(nb, ext={"timestep": steps, "highlevel": nb["highlevel_dense"], "noise": noised}, get_feature=True)
    est = est.reshape(noise.shape)
    if getattr(args, "stl_bc_mask", False):
        # reference :435-437 with dense_scores = tj_scores_prior (:1282-1283): rows of non-existent lanes and traj-opt
        # samples that violate the spec carry no denoising loss (the parser forces stl_bc_mask, :1781)
        if "tj_scores_prior" not in nb:
            raise KeyError("train_step_ddpm with stl_bc_mask needs batch['tj_scores_prior'] (bs, n_randoms, 3)")
        n = bs * S * 3
        m = (nb["tj_scores_prior"].reshape(n, 1) * nb["valids_dense"].reshape(n, 1) > 0).float()
        loss_diffusion = mask_mean(torch.square(noise - est), m)
    else:
        loss_diffusion = torch.mean(torch.square(noise - est))
    rd = {"loss_diffusion": loss_diffusion, "est_cmds_a": est, "feature": feature}
    rd["loss_stl"] = rd["loss_diffusion"].detach() * 0
    rd["loss"] = rd["loss_diffusion"]
    if optimizer is not None:
        optimizer.zero_grad()
        rd["loss"].backward()
        optimizer.step()
    return rd


def train_step_rect(net, batch_cuda, stls_cac, coeffs, args, optimizer=None, gt_stlp=None):
    """One training iteration of the --rect_head stage (reference nusc_train.py:1352-1427, 1523-1525; README "Ours"
    training command): chains sampled with the frozen denoiser, best of the last ``multi_cands`` iterates,
    ``rect_forward`` on the detached controls and scores, rollout, ``compute_policy_loss``, backward into rect_net and
    the optimiser step (``torch.optim.Adam(net.rect_net.parameters())`` upstream).  pSTL parameters per chain from
    ``pre_stlp`` (training layout); ``gt_stlp`` (bs, 6) only feeds the ground-truth scores of the report.  The
    denoising loss upstream logs next to it (not part of this stage's ``loss``) is not evaluated.  Returns ``rd``."""
    S = args.n_randoms
    bs = batch_cuda["ego_traj"].shape[0]
    N = bs * S * 3
    nb = LazyBatch({k: batch_cuda[k] for k in ("ego_traj", "neighbors", "currlane_wpts", "leftlane_wpts", "rightlane_wpts",
                                               "curr_id", "left_id", "right_id", "gt_high_level", "pre_stlp")
                    if k in batch_cuda})
    nb["neighbor_trajs_aug"] = batch_cuda["neighbors_traj"][..., :7]
    if gt_stlp is None:
        gt_stlp = batch_cuda["pre_stlp"].reshape(bs, -1, 3, 6)[:, 0, 0]
    nb = augment_batch_data(nb, gt_stlp, args)
    pack = nb["_pstl_pack"]
    hl = nb["highlevel_dense"]
    progs = _fused_programs(stls_cac, args.nt)
    with torch.no_grad():
        noise = torch.empty((N, args.nt * 2), device=hl.device)
        res = diffusion_rollout(noise, net, nb, hl, None, args, coeffs, n_randoms=S, return_feature=True,
                                scene_feature_only=True)
        nn_controls, feature = res[0], res[1]
        if args.multi_cands is not None:
            r = score_pack(pack, res[2].stacked_last(args.multi_cands), args, progs, want=("best_score", "best_controls"))
            nn_controls, prev_scores = r["best_controls"], r["best_score"]
        else:
            prev_scores = score_pack(pack, nn_controls, args, progs)["best_score"]
        nn_trajs = generate_trajs(pack.state0, nn_controls, args.dt)
    rect_controls = net.rect_forward(feature, hl, nb["stlp_dense"][:, 0], nn_controls.detach(), prev_scores.detach())
    rect_trajs = generate_trajs(pack.state0, rect_controls, args.dt)
    zeros = torch.zeros((N, args.nt * 2), device=hl.device)
    extras = (None, zeros, hl, prev_scores, nb["valids_dense"].reshape(-1), 0, zeros, nn_controls, None, rect_controls)
    rd, all_scores = compute_policy_loss(nb, None, stls_cac, nn_trajs, rect_trajs, None, args, diffusion_extras=extras)
    rd["all_scores"] = all_scores
    if optimizer is not None:
        optimizer.zero_grad()
        rd["loss"].backward()
        optimizer.step()
    return rd


def trajopt(batch_cuda, stls_cac, args, iters=None, params=None, record=None, state=None):
    """Trajectory optimisation of the stored control parameters (reference nusc_train.py:287-316, 1303-1325):
    ``iters`` (default ``args.traj_opt_iters``) Adam steps, lr ``args.trajopt_lr``, on
    ``mean(relu(stl_trajopt_thres - score) * valid) / clip(mean(valid), 1e-3) + reg_loss * (mean relu(w^2 - w_max^2) + ...)``.
    ``batch_cuda`` is a scene batch after ``augment_batch_data`` (training layout: ``stlp_dense`` from ``pre_stlp``);
    ``params`` defaults to ``batch_cuda["params"]`` (bs, n_randoms, 3, nt, 2).  Every iteration is two launches: the
    fused rollout + STL reverse-mode kernel and the regulariser + Adam update (upstream: ~600 autograd launches).
    Returns (optimised params, same shape; scores (N,) of the iterate BEFORE the last step, as upstream logs them).
    ``record(ii, scores)`` is called after every iteration when given (forces no sync by itself).
    ``state`` = (adam_m, adam_v, first_iteration) resumes a run (Adam moments (N, nt, 2), updated in place)."""
    _check_fused_supported(args, "trajopt")
    iters = int(args.traj_opt_iters if iters is None else iters)
    p0 = batch_cuda["params"] if params is None else params
    _nv.require_cuda(p0, "params")
    S, nt = args.n_randoms, args.nt
    bs = batch_cuda["currlane_wpts"].shape[0]
    N = bs * S * 3
    pack = batch_cuda.get("_pstl_pack")
    if pack is None:
        pack = ScenePack.from_batch(batch_cuda, batch_cuda["stlp_dense"], S)
    progs = _fused_programs(stls_cac, nt)
    if progs is None:
        raise NotImplementedError("trajopt needs the typed spec of build_stl_cache")
    p = _nv.f32(p0.reshape(N, nt, 2)).clone()
    it0 = 0
    if state is None:
        m, v = torch.zeros_like(p), torch.zeros_like(p)
    else:
        m, v, it0 = _nv.f32(state[0].reshape(N, nt, 2)), _nv.f32(state[1].reshape(N, nt, 2)), int(state[2])
    scores = torch.empty((N,), dtype=torch.float32, device=p.device)
    L = _nv.lib()
    pa = _nv.prog_array(progs)
    sv, sp = pack.view(), _spec(args)
    ws = _nv.workspace(N * nt * 2 * 4 + L.pstl_score_workspace_bytes(pa, N, nt, 1), p.device, "trajopt")
    # one scalar per call, computed on the device side of the host API (valid is a batch constant)
    inv_norm = 1.0 / (N * max(float(pack.valid.mean().item()), 1e-3))
    for ii in range(it0, it0 + iters):
        _nv.check(L.pstl_trajopt_step(pa, _nv.C.byref(sv), _nv.C.byref(sp), _nv.fptr(pack.mode), _nv.fptr(pack.state0),
                                      _nv.fptr(pack.stlp), _nv.fptr(pack.valid), N, _nv.C.c_float(args.stl_trajopt_thres),
                                      _nv.C.c_float(inv_norm), _nv.C.c_float(args.reg_loss), _nv.C.c_float(args.mul_w_max),
                                      _nv.C.c_float(args.mul_a_max), _nv.C.c_float(args.trajopt_lr), ii, _nv.fptr(p),
                                      _nv.fptr(m), _nv.fptr(v), _nv.fptr(scores), _nv.ptr(ws), _nv.stream()),
                  "pstl_trajopt_step")
        if record is not None:
            record(ii, scores)
    return p.reshape(p0.shape), scores


def save_trajopt_params(params, iter_i, traj_i, ti, args, save_stlp=None, model_dir=None):
    """The traj-opt outputs in the reference's on-disk format (nusc_train.py:775-797): one ``.npy`` per scene sample,
    ``params_<traj>_<ti>.npy`` ("final"), ``..._init.npy`` ("init"), ``..._iter<k>.npy`` (integer), ``scores_<traj>_<ti>.npy``
    ("scores"), plus ``..._stlp.npy`` (n_randoms, 3, 1, 6) when ``save_stlp`` is given.  Nothing is written under --test."""
    if getattr(args, "test", False):
        return []
    model_dir = model_dir or args.model_dir
    os.makedirs(model_dir, exist_ok=True)
    bs = params.shape[0]
    arr = params.detach().cpu().numpy()
    stlp = None
    if save_stlp is not None:
        stlp = save_stlp.detach().cpu().numpy().reshape(bs, args.n_randoms, 3, 1, save_stlp.shape[-1])
    written = []
    for i in range(bs):
        key = (int(traj_i[i]), int(ti[i]))
        if iter_i == "scores":
            name = "scores_%05d_%04d.npy" % key
        elif iter_i == "init":
            name = "params_%05d_%04d_init.npy" % key
        elif iter_i == "final":
            name = "params_%05d_%04d.npy" % key
        else:
            name = "params_%05d_%04d_iter%05d.npy" % (key + (int(iter_i),))
        np.save(os.path.join(model_dir, name), arr[i])
        written.append(name)
        if stlp is not None:
            name = "params_%05d_%04d_stlp.npy" % key
            np.save(os.path.join(model_dir, name), stlp[i])
            written.append(name)
    return written


def load_trajopt_params(model_dir, traj_i, ti, load_stlp=True):
    """What the reference's dataset reads back per sample (nusc_dataset.py:203-225): ``params`` / ``params_init``
    (n_randoms, 3, nt, 2) and, with --load_stlp, ``pre_stlp`` (n_randoms, 3, 1, 6) and ``tj_scores_prior`` (n_randoms, 3),
    stacked over the given (traj_i, ti) pairs."""
    out = {"params": [], "params_init": []}
    if load_stlp:
        out.update(pre_stlp=[], tj_scores_prior=[])
    for a, b in zip(traj_i, ti):
        key = (int(a), int(b))
        out["params"].append(np.load(os.path.join(model_dir, "params_%05d_%04d.npy" % key)))
        out["params_init"].append(np.load(os.path.join(model_dir, "params_%05d_%04d_init.npy" % key)))
        if load_stlp:
            out["pre_stlp"].append(np.load(os.path.join(model_dir, "params_%05d_%04d_stlp.npy" % key)))
            out["tj_scores_prior"].append(np.load(os.path.join(model_dir, "scores_%05d_%04d.npy" % key)))
    return {k: torch.from_numpy(np.stack(v)).float() for k, v in out.items()}


def closed_loop_pick(scores_all, ego_controls, ego_trajs):
    """candidate selection of the closed-loop simulator (reference nusc_sim.py:677-683): chains are rows
    ``n = sample*3 + mode`` of ONE scene; modes 1, 2 (lane changes) are masked to -1e4 and the global arg-max wins
    (first maximum).  Returns (index into the n rows, highest score, controls (1,T,2), trajectory (1,T+1,4)).
    The mask is applied to a copy (upstream overwrites its score tensor in place)."""
    n = scores_all.shape[0]
    cube = scores_all.reshape(n // 3, 3).clone()
    cube[:, 1:3] = -10000
    total_idx = torch.argmax(cube)
    return total_idx, cube.flatten()[total_idx], ego_controls[total_idx].unsqueeze(0), ego_trajs[total_idx].unsqueeze(0)


class CapturedPipeline:
    """``sample_and_score`` captured ONCE into a CUDA graph and replayed per batch (static shapes): the
    ~170 kernel launches of a batch become one ``cudaGraphLaunch``, so the GPU is no longer paced by the
    Python/ctypes launch path.  Same kernels, same arithmetic as the eager call.

    ``runner = CapturedPipeline(net, stls, coeffs, args, example_batch); out = runner(batch)``:
    ``batch`` tensors (host-pinned or device) are copied into the graph's static inputs, the graph is
    replayed, and ``out`` holds the graph's static output tensors (valid until the next call).
    Noise: x_T and the z stream come from the sampler's Philox counter plus a device word that the graph bumps on every replay
    (``pstl_denoiser_set_noise_counter``), so replays draw fresh normals as upstream's ``randn_like`` does.
    ``--guidance`` is captured too: its batch normaliser lives in a device word (``pstl_guidance_cfg.inv_norm_dev``).
    Not capturable: injected noise (the test mode) and ``--refinement`` (a host-driven optimiser loop)."""

    KEYS = ("ego_traj", "neighbors", "neighbors_traj", "currlane_wpts", "leftlane_wpts", "rightlane_wpts", "curr_id",
            "left_id", "right_id", "gt_high_level", "pre_stlp")

    def __init__(self, net, stls_cac, coeffs, args, example_batch, warmup=2, counter_start=0, counter_stride=4096):
        if getattr(args, "inject_noise", None) is not None or getattr(args, "refinement", False):
            raise NotImplementedError("CapturedPipeline: injected noise / --refinement run on the eager path")
        self.net, self.stls, self.coeffs, self.args = net, stls_cac, coeffs, args
        dev = next(net.parameters()).device
        _nv.require_cuda(next(net.parameters()), "model parameters")
        self.static_in = {k: example_batch[k].to(dev, copy=True) for k in self.KEYS if k in example_batch}
        # Philox step words of replay r: counter_start + r * counter_stride + (1 .. diffusion steps); BatchPipeliner deals its
        # runners interleaved ranges
        self.counter = torch.full((1,), int(counter_start), dtype=torch.int64, device=dev)
        self._counter_stride = int(counter_stride)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # allocator / lazy caches / workspaces settle before the capture
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # a capture stream of its own: the library's scratch buffers are keyed by (device, tag, stream), so two runners
        # never share workspace and may replay concurrently on two streams (torch's default capture stream is shared)
        self._capture_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.graph(self.graph, stream=self._capture_stream):
            self.out = self._body()
        self._weights_key = net._weights_key()
        self._handle_value = net.native_handle(getattr(args, "precision", "fp32")).value
        # double buffering of the inputs: `prefetch` fills a staging copy on its own stream while a replay runs
        self.staging = {k: torch.empty_like(t) for k, t in self.static_in.items()}
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._staged = torch.cuda.Event()
        self._staging_free = torch.cuda.Event()
        self._staging_free.record(torch.cuda.current_stream(dev))
        self._have_prefetch = False

    def prefetch(self, batch):
        """Start copying the NEXT batch (pinned host or device tensors) into the staging inputs on a separate stream;
        the copy overlaps whatever the main stream is running.  The next ``runner()`` call consumes it."""
        self._check(batch)
        self._copy_stream.wait_event(self._staging_free)
        with torch.cuda.stream(self._copy_stream):
            for k, t in self.staging.items():
                t.copy_(batch[k], non_blocking=True)
            self._staged.record(self._copy_stream)
        self._have_prefetch = True

    def _body(self):
        handle = self.net.native_handle(getattr(self.args, "precision", "fp32"))
        _nv.check(_nv.lib().pstl_denoiser_set_noise_counter(handle, _nv.C.c_void_p(self.counter.data_ptr())),
                  "pstl_denoiser_set_noise_counter")
        self.counter.add_(self._counter_stride)  # > diffusion steps: replays use disjoint Philox step words
        with torch.no_grad():
            return sample_and_score(self.net, self.static_in, self.stls, self.coeffs, self.args)

    def __call__(self, batch=None):
        cur = torch.cuda.current_stream()
        if batch is None:
            if not self._have_prefetch:
                raise ValueError("CapturedPipeline(): no batch given and nothing prefetched")
            cur.wait_event(self._staged)
            for k, t in self.static_in.items():
                t.copy_(self.staging[k], non_blocking=True)  # device to device, a few microseconds
            self._staging_free.record(cur)
            self._have_prefetch = False
        else:
            self._check(batch)
            for k, t in self.static_in.items():
                if batch[k] is not t:
                    t.copy_(batch[k], non_blocking=True)
        if self.net._weights_key() != self._weights_key:
            # in-place parameter updates: the handle re-derives its weight copies in place on this stream, ahead of the
            # replay; parameters that moved to other storage rebuilt the handle the graph baked in
            if self.net.native_handle(getattr(self.args, "precision", "fp32")).value != self._handle_value:
                raise RuntimeError("CapturedPipeline: the model's parameters moved since the capture; build a new runner")
            self._weights_key = self.net._weights_key()
        self.graph.replay()
        return self.out

    def _check(self, batch):
        """a captured graph has static shapes: refuse a batch of another shape instead of broadcasting it silently"""
        for k, t in self.static_in.items():
            if k not in batch or tuple(batch[k].shape) != tuple(t.shape):
                raise ValueError("CapturedPipeline was captured for %s%s, got %s" %
                                 (k, tuple(t.shape), tuple(batch[k].shape) if k in batch else "nothing"))


class BatchPipeliner:
    """``depth`` CapturedPipeline runners replayed round-robin, each on its own CUDA stream, so CONSECUTIVE BATCHES OVERLAP: a
    batch's pipeline is a chain of kernels that each leave SMs idle at some point (the persistent sampler's last partial
    wave — 1,536 tiles over 148 SMs —, the scorer's last wave, the latency-bound encoder / RefineNet front-end kernels), and
    the next batch's kernels run there.  Same kernels and arithmetic per batch; measured at BASELINE config 2: 5.15 -> 4.67 ms
    per batch with depth 2.  Every runner owns its inputs, outputs, workspaces (the library's scratch is keyed by capture
    stream) and an interleaved range of Philox step words.

    ``out, done, slot = pipe.submit(batch)``: ``batch`` (pinned host or device tensors) is copied into the runner's inputs on
    the runner's stream, which first waits for the caller's current stream; ``out`` holds that runner's static outputs,
    valid from ``done`` (a CUDA event on ``pipe.streams[slot]``) until the runner is used again ``depth`` submits later —
    enqueue whatever reads them on ``pipe.streams[slot]`` or after ``done``.  ``drain()`` makes the current stream wait for
    everything submitted.  The runners share the model's weight images read-only: ``drain()`` and synchronise before an
    optimiser step or ``load_state_dict`` (an in-place refresh would race the other runner's replay)."""

    def __init__(self, net, stls_cac, coeffs, args, example_batch, depth=2):
        dev = next(net.parameters()).device
        self.depth = int(depth)
        self.runners = [CapturedPipeline(net, stls_cac, coeffs, args, example_batch, counter_start=k * 4096,
                                         counter_stride=self.depth * 4096) for k in range(self.depth)]
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(self.depth)]
        self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self._n = 0

    def submit(self, batch):
        k = self._n % self.depth
        self._n += 1
        st = self.streams[k]
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            out = self.runners[k](batch)
            self.done[k].record(st)
        return out, self.done[k], k

    def drain(self):
        cur = torch.cuda.current_stream()
        for st in self.streams:
            cur.wait_stream(st)


def run_sampling_test(stls_cac, data_loader, net, coeffs, args, result_queue=None, thread_nusc=None):
    """open-loop sampling test over ``data_loader`` (any iterable of batch dicts): the timed region, then the metrics of
    the reference's report line (acc, scene_acc, ade, fde, std, vol; area / entropies with --run_sampling_test's
    extra_diversity) computed on the device (``pstl_b200.metrics``).  Visualisation is out of scope."""
    from . import metrics as M
    meters = {}
    results = []
    for bi, batch in enumerate(data_loader):
        if bi > args.n_trials:
            continue
        batch_cuda = {k: (v.cuda() if hasattr(v, "cuda") else v) for k, v in batch.items()}
        torch.cuda.synchronize()
        t1 = time.time()
        out = sample_and_score(net, batch_cuda, stls_cac, coeffs, args)
        torch.cuda.synchronize()
        t2 = time.time()
        S, nt = args.sampling_size, args.nt
        bs = batch_cuda["ego_traj"].shape[0]
        trajs, valid = out["trajs"], out["pack"].valid.reshape(bs, S, 3)
        ma_std, ma_vol, _, _ = M.measure_diversity(trajs[:, :-1, :2].reshape(bs, S, 3, nt * 2), out["scores"].reshape(bs, S, 3),
                                                   valid, nt)
        ade, fde = M.compute_ade_fde(batch_cuda["ego_traj"][..., :4], trajs[:, :-1, :4], valid)
        vals = [("acc", out["acc"].item()), ("scene_acc", out["scene_acc"].item()), ("ade", ade.item()), ("fde", fde.item()),
                ("std", float(ma_std)), ("vol", float(ma_vol)), ("time", t2 - t1)]
        if getattr(args, "extra_diversity", False):
            ex = M.measure_extra_diversity(trajs[:, :-1].reshape(bs, S, 3, nt * 4), out["scores"].reshape(bs, S, 3), valid, nt,
                                           out["controls"].reshape(bs, S, 3, nt * 2), -args.mul_w_max, args.mul_w_max,
                                           -args.mul_a_max, args.mul_a_max)
            vals += [(k, v.item()) for k, v in ex.items()]
        for k, v in vals:
            meters.setdefault(k, []).append(v)
        mean = lambda k: float(np.mean(meters[k])) if k in meters else float("nan")
        print("###[%02d] NN acc:%.3f scene_acc:%.3f ade:%.3f fde:%.3f std:%.3f vol:%.3f area:%.3f s:%.3f u:%.3f ||| T:%.3f"
              % (bi, mean("acc"), mean("scene_acc"), mean("ade"), mean("fde"), mean("std"), mean("vol"), mean("area"),
                 mean("ent_s"), mean("ent_wa"), mean("time")))
        results.append(out)
    return meters, results


# ---------------------------------------------------------------------------------------
# flags (reference :1635-1814)
# ---------------------------------------------------------------------------------------

def generate_parser(argv=None):
    """Same flags and post-parse overrides as the reference parser, plus --synthetic / --precision."""
    parser = argparse.ArgumentParser("")
    add = parser.add_argument
    add("--seed", type=int, default=1007)
    add("--exp_name", "-e", type=str, default=None)
    add("--gpus", type=str, default="0")
    add("--epochs", type=int, default=500)
    add("--test", action="store_true", default=False)
    add("--net_pretrained_path", "-P", type=str, default=None)
    add("--num_workers", type=int, default=8)
    add("--batch_size", "-b", type=int, default=128)
    add("--lr", type=float, default=3e-4)
    add("--hiddens", type=int, nargs="+", default=[256, 256])
    for name, dflt in (("print_freq", 10), ("save_freq", 100), ("viz_freq", 50), ("num_viz", 10)):
        add("--" + name, type=int, default=dflt)
    for name in ("no_viz", "mini", "collect_data", "offline", "refined_safety", "debug", "use_gt_stlp", "skip_nusc_load",
                 "clip_dist", "gt_nei", "stl_bc_mask", "trajopt_only", "inline", "use_init_hint",
                 "generate_split_on_the_fly", "check_stl_params", "norm_stl", "flex", "load_stlp", "load_tj", "bc", "vae",
                 "diffusion", "cos", "grad_rollout", "rect_head", "joint", "not_use_rect", "measure_diversity",
                 "extra_diversity", "viz_correct", "run_sampling_test", "replace_hint", "diff_full", "refinement",
                 "raw_refinement", "diverse_loss", "no_arch", "diverse_detach", "test_t1", "test_scenes",
                 "test_aggressive", "viz_last", "lite_refine", "interval", "diffusion_clip", "gt_data_training",
                 "guidance", "guidance_reverse", "oracle_filter", "clip_rect", "ego", "other", "backup", "no_refinenet",
                 "time_profile"):
        add("--" + name, action="store_true", default=False)
    add("--train_ratio", type=float, default=0.7)
    add("--n_neighbors", "-N", type=int, default=8)
    add("--n_randoms", type=int, default=64)
    add("--n_segs", type=int, default=15)
    add("--n_expands", type=int, default=4)
    add("--cache_path", type=str, default="e0_nusc_cache")
    add("--ego_L", type=float, default=4.084)
    add("--ego_W", type=float, default=1.730)
    add("--refined_nL", type=int, default=4)
    add("--refined_nW", type=int, default=1)
    add("--nt", type=int, default=20)
    add("--dt", type=float, default=0.5)
    add("--mul_w_max", type=float, default=0.5)
    add("--mul_a_max", type=float, default=5.0)
    add("--smoothing_factor", type=float, default=100.0)
    add("--anno_path", type=str, default="annotated_data_trainval")
    add("--stl_nn_thres", type=float, default=0.0005)
    add("--stl_trajopt_thres", type=float, default=0.01)
    add("--traj_opt_iters", type=int, default=2000)
    add("--trajopt_lr", type=float, default=0.005)
    add("--opt_epochs", type=int, default=0)
    add("--trajopt_save_freq", type=int, default=1000)
    add("--params_load_path", "-P2", type=str, default="e1_nusc_trajopt")
    add("--filter_traj", type=int, nargs="+", default=None)
    add("--stl_weight", type=float, default=1.0)
    add("--bc_weight", type=float, default=0.0)
    add("--vae_dim", type=int, default=64)
    add("--weight_vae_bc", type=float, default=1.0)
    add("--weight_vae_kl", type=float, default=1.0)
    add("--diffusion_steps", type=int, default=100)
    add("--diffusion_weight", type=float, default=1.0)
    add("--beta_start", type=float, default=1e-4)
    add("--beta_end", type=float, default=0.02)
    add("--reg_loss", type=float, default=10.0)
    add("--rect_hiddens", type=int, nargs="+", default=[256, 256])
    add("--rect_reg_loss", type=float, default=0.000)
    add("--extra_rect_reg", type=float, default=0.0)
    add("--epi_print_freq", type=int, default=1)
    add("--sampling_size", type=int, default=64)
    add("--n_trials", type=int, default=100)
    add("--diversity_weight", type=float, default=1.0)
    add("--diversity_scale", type=float, default=1.0)
    add("--n_shards", type=int, default=4)
    add("--diverse_fuse_type", type=str, default="add")
    add("--multi_cands", type=int, default=None)
    add("--collision_loss", type=float, default=None)
    add("--guidance_niters", type=int, default=3)
    add("--guidance_before", type=int, default=1000)
    add("--guidance_lr", type=float, default=0.01)
    add("--guidance_sets", nargs="+", type=int, default=None)
    add("--guidance_freq", type=int, default=None)
    add("--n_rolls", type=int, default=None)
    add("--suffix", type=str, default=None)
    # additive flags (not in the reference)
    add("--synthetic", type=int, default=None, help="run on this many synthetic scenes per batch")
    add("--precision", type=str, default="fp32", choices=["fp32", "bf16", "f16", "f16x3"],
        help="denoiser arithmetic: fp32 SIMT, bf16 / fp16 tensor-core operands (2e-2 bound; fp16 8x closer), or split fp16 operands (fp32-grade, 1e-5)")
    args = parser.parse_args(argv)
    # post-parse overrides, reference :1780-1812
    args.gt_nei = True
    args.stl_bc_mask = True
    args.cos = True
    if not args.collect_data and not args.trajopt_only:
        args.measure_diversity = True
    if args.run_sampling_test:
        args.test = True
        args.extra_diversity = True
    if args.collect_data:
        args.epochs, args.batch_size, args.viz_freq, args.print_freq, args.uturn = 1, 1024, 10, 1, True
    if args.trajopt_only:
        args.opt_epochs, args.epochs, args.batch_size, args.viz_freq = 1, 1, 1024, 10
        args.diffusion, args.num_viz, args.flex = True, 256, True
    if args.opt_epochs > 0:
        args.epochs = args.opt_epochs
    if args.load_stlp:
        args.load_tj = True
    if args.rect_head:
        args.interval = True
        args.diffusion_clip = True
        args.diff_full = True
    args.offline = not args.collect_data
    if args.test:
        args.epochs = 1
    return args


OURS_FLAGS = ["-e", "e7_ours", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--rect_head", "--flex",
              "--diverse_loss", "--multi_cands", "5", "--test", "--run_sampling_test", "--skip_nusc_load",
              "--viz_correct"]
GUIDANCE_FLAGS = ["-e", "e7_ours", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--rect_head", "--flex",
                  "--diverse_loss", "--multi_cands", "10", "--test", "--run_sampling_test", "--viz_correct",
                  "--guidance", "--guidance_before", "10", "--guidance_niters", "1", "--guidance_lr", "0.01",
                  "--n_rolls", "3", "--other", "--skip_nusc_load"]


def default_args(flags=None, **over):
    """parsed defaults (README "Ours" flags unless given) with attribute overrides — for tests/bench."""
    args = generate_parser(list(OURS_FLAGS if flags is None else flags))
    for k, v in over.items():
        setattr(args, k, v)
    return args


def save_model_freq_last(state_dict, model_dir, epi, save_freq, epochs):
    """checkpoint cadence of the reference (utils.py:81-85): ``model_%05d.ckpt`` every ``save_freq`` epochs and at the
    end, ``model_last.ckpt`` every 10 epochs and at the end; plain ``torch.save`` of the state_dict."""
    os.makedirs(model_dir, exist_ok=True)
    if epi % save_freq == 0 or epi == epochs - 1:
        torch.save(state_dict, "%s/model_%05d.ckpt" % (model_dir, epi))
    if epi % 10 == 0 or epi == epochs - 1:
        torch.save(state_dict, "%s/model_last.ckpt" % (model_dir))


def run_training(stls_cac, train_loader, net, coeffs, args, model_dir=None, log=print):
    """The epoch loop of the reference's two diffusion stages (nusc_train.py:1228-1235 optimiser, :1245-1577 loop, train
    mode only): README step 1 (``--diffusion``: ``train_step_ddpm``, Adam over every parameter) or step 2
    (``--rect_head``: ``train_step_rect``, Adam over rect_net).  ``args.epochs`` passes over ``train_loader`` (scene
    batches on the host or the device), the per-epoch means of the logged terms, checkpoints in upstream's layout when
    ``model_dir`` is given.  Returns the list of per-epoch dicts."""
    if not args.diffusion or getattr(args, "joint", False):
        raise NotImplementedError("only the --diffusion stages (denoiser; --rect_head RefineNet) are built")
    if args.rect_head:
        optimizer = torch.optim.Adam(net.rect_net.parameters(), lr=args.lr)
        step = lambda b: train_step_rect(net, b, stls_cac, coeffs, args, optimizer)
    else:
        optimizer = torch.optim.Adam(net.parameters(), lr=args.lr)
        step = lambda b: train_step_ddpm(net, b, coeffs, args, optimizer)
    keys = ("loss", "loss_diffusion", "loss_stl", "loss_reg", "loss_diversity", "extra_loss_reg", "acc", "avg_speed")
    history = []
    for epi in range(args.epochs):
        sums = {}
        nb = 0
        for batch in train_loader:
            batch_cuda = {k: (v.cuda(non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
            rd = step(batch_cuda)
            for k in keys:
                if k in rd:
                    sums[k] = sums.get(k, 0.0) + rd[k].detach()
            nb += 1
        means = {k: float(v) / max(nb, 1) for k, v in sums.items()}  # one device->host read per term and epoch
        history.append(means)
        if epi % max(int(getattr(args, "epi_print_freq", 1)), 1) == 0 or epi == args.epochs - 1:
            log("[%05d/%05d] " % (epi, args.epochs) + " ".join("%s:%.4f" % (k, means[k]) for k in keys if k in means))
        if model_dir is not None:
            save_model_freq_last(net.state_dict(), model_dir, epi, args.save_freq, args.epochs)
    return history


run_rect_training = run_training


def main(argv=None):
    """``python -m pstl_b200.nusc_train ... --run_sampling_test --synthetic 32`` (open-loop test) or, without
    ``--run_sampling_test``, the README's training stages on synthetic scene batches: step 1 (``--diffusion``, the
    denoiser) or step 2 (``--rect_head``, RefineNet); ``--synthetic B`` scenes per batch, ``--epochs``, checkpoints
    under ``./exps_nusc/<exp_name>/models``."""
    from . import synthetic
    from .nusc_model import Net
    args = generate_parser(argv)
    if not args.run_sampling_test:
        if not args.diffusion or getattr(args, "trajopt_only", False):
            raise SystemExit("training: the --diffusion stages are built (denoiser, --rect_head); see nusc_train.trajopt "
                             "for the trajectory optimisation")
        torch.manual_seed(args.seed)
        net = Net(args).cuda()
        if args.net_pretrained_path is not None:
            net.load_state_dict(torch.load(args.net_pretrained_path), strict=False)
        if args.synthetic is None and os.path.isfile(args.cache_path):
            # real data, offline: --cache_path <cache.npz written by --collect_data>, traj-opt files under
            # --params_load_path (a directory holding params_*.npy), optional split file in PSTL_SPLIT_FILE
            from . import nusc_dataset
            pdir = args.params_load_path if os.path.isdir(args.params_load_path or "") else None
            loader = nusc_dataset.get_dataloader(args, args.cache_path, os.environ.get("PSTL_SPLIT_FILE"), pdir)
        else:
            bs = args.synthetic or args.batch_size
            loader = [synthetic.make_scene_batch(bs, nt=args.nt, dt=args.dt, n_neighbors=args.n_neighbors, n_segs=args.n_segs,
                                                 n_randoms=args.n_randoms, seed=args.seed + i) for i in range(3)]
        model_dir = os.path.join("exps_nusc", args.exp_name or "rect", "models")
        return run_training(build_stl_cache(args), loader, net, get_diffusion_coeffs(args), args, model_dir)
    torch.manual_seed(args.seed)
    stls_cac = build_stl_cache(args)
    net = Net(args).cuda()
    if args.net_pretrained_path is not None:
        net.load_state_dict(torch.load(args.net_pretrained_path), strict=(not args.rect_head))
    coeffs = get_diffusion_coeffs(args)
    if args.synthetic is None and os.path.isfile(args.cache_path):  # offline real data, see the training branch
        from . import nusc_dataset
        pdir = args.params_load_path if os.path.isdir(args.params_load_path or "") else None
        loader = nusc_dataset.get_dataloader(args, args.cache_path, os.environ.get("PSTL_SPLIT_FILE"), pdir, shuffle=False)
    else:
        bs = args.synthetic or args.batch_size
        loader = [synthetic.make_scene_batch(bs, nt=args.nt, dt=args.dt, n_neighbors=args.n_neighbors,
                                             n_segs=args.n_segs, n_randoms=args.n_randoms, seed=args.seed + i)
                  for i in range(min(args.n_trials + 1, 3))]
    return run_sampling_test(stls_cac, loader, net, coeffs, args)


if __name__ == "__main__":
    main()
