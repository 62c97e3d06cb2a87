"""Seeded synthetic scene batches of the NuScenes cache shape.

NuScenes is unavailable offline, so the benchmark and the parity tests run on
synthetic scenes with the same keys/shapes the reference dataset yields
(reference nusc_dataset.py:227-232): ego state, three 15-point lane centre-lines,
``Knei`` neighbour vehicles with constant-velocity tracks and the six pSTL
parameters (vmin, vmax, dmin, dmax, dsafe, thmax) per (scene, sample, mode).
Everything is fp32 on the CPU; callers move it to the GPU.
"""
import math

import torch

EGO_L = 4.084
EGO_W = 1.730


def _u(gen, lo, hi, *size):
    return torch.rand(*size, generator=gen) * (hi - lo) + lo


def make_scene_batch(bs, nt=20, dt=0.5, n_neighbors=8, n_segs=15, n_randoms=64, seed=1007,
                     lane_valid_p=0.7, nei_valid_p=0.7):
    """Return a dict batch with the keys the reference loader produces.

    ego_traj (bs,nt,6) neighbors (bs,K,7) neighbors_traj (bs,K,nt,7)
    {curr,left,right}lane_wpts (bs,nseg,3) {curr,left,right}_id (bs,1)
    gt_high_level (bs,1) pre_stlp (bs,R,3,1,6) params,params_init (bs,R,3,nt,2)
    tj_scores_prior (bs,R,3) traj_i,ti (bs,)
    """
    g = torch.Generator().manual_seed(int(seed))
    ex = _u(g, -50, 50, bs)
    ey = _u(g, -50, 50, bs)
    eth = _u(g, -0.1, 0.1, bs)
    ev = _u(g, 2, 8, bs)

    # ground-truth ego track: mild random controls through the Euler unicycle
    gw = _u(g, -0.05, 0.05, bs, nt) * 0.3
    ga = _u(g, -1, 1, bs, nt)
    st = torch.stack([ex, ey, eth, ev], -1)
    track = [st]
    for t in range(nt - 1):
        x, y, th, v = track[-1].unbind(-1)
        nxt = torch.stack([x + v * torch.cos(th) * dt, y + v * torch.sin(th) * dt,
                           th + gw[:, t] * dt, v + ga[:, t] * dt], -1)
        track.append(nxt)
    track = torch.stack(track, 1)  # (bs,nt,4)
    ego_traj = torch.cat([track, torch.full((bs, nt, 1), EGO_L), torch.full((bs, nt, 1), EGO_W)], -1)

    # lanes: arc-length samples along the heading with a small curvature term
    s = torch.linspace(-5.0, 100.0, n_segs)[None, :]  # (1,nseg)
    curv = _u(g, -2e-3, 2e-3, bs)[:, None]
    lanes, ids = {}, {}
    for name, off in (("curr", 0.0), ("left", 3.5), ("right", -3.5)):
        # ego-frame polyline: x=s, y=off+0.5*curv*s^2, heading=atan(curv*s)
        lx = s.expand(bs, -1)
        ly = off + 0.5 * curv * s * s
        lth = torch.atan(curv * s)
        c, sn = torch.cos(eth)[:, None], torch.sin(eth)[:, None]
        wx = ex[:, None] + lx * c - ly * sn
        wy = ey[:, None] + lx * sn + ly * c
        wth = lth + eth[:, None]
        lane = torch.stack([wx, wy, wth], -1)
        if name == "curr":
            valid = torch.ones(bs, 1)
        else:
            valid = (torch.rand(bs, 1, generator=g) < lane_valid_p).float()
        lanes[name] = lane * valid[:, :, None]  # invalid lanes are zero-filled
        ids[name] = valid

    # neighbours: ego-frame box, constant-velocity tracks, invalid rows all-zero
    K = n_neighbors
    nvalid = (torch.rand(bs, K, generator=g) < nei_valid_p).float()
    rx = _u(g, -30, 60, bs, K)
    ry = _u(g, -8, 8, bs, K)
    # keep neighbours off the ego's own footprint at t=0
    ry = torch.where((rx.abs() < 8) & (ry.abs() < 2.5), ry.sign() * 3.0 + ry, ry)
    nth = eth[:, None] + _u(g, -0.2, 0.2, bs, K)
    nv = _u(g, 0, 8, bs, K)
    nL = _u(g, 3.5, 5.5, bs, K)
    nW = _u(g, 1.6, 2.2, bs, K)
    c, sn = torch.cos(eth)[:, None], torch.sin(eth)[:, None]
    nx = ex[:, None] + rx * c - ry * sn
    ny = ey[:, None] + rx * sn + ry * c
    tt = (torch.arange(nt).float() * dt)[None, None, :]
    tx = nx[..., None] + nv[..., None] * torch.cos(nth)[..., None] * tt
    ty = ny[..., None] + nv[..., None] * torch.sin(nth)[..., None] * tt
    ones = torch.ones(bs, K, nt)
    ntraj = torch.stack([ones, tx, ty, nth[..., None] * ones, nv[..., None] * ones,
                         nL[..., None] * ones, nW[..., None] * ones], -1)
    ntraj = ntraj * nvalid[:, :, None, None]
    neighbors = ntraj[:, :, 0, :].clone()

    R = n_randoms
    stlp = torch.stack([_u(g, 0, 2, bs, R, 3), _u(g, 8, 12, bs, R, 3), _u(g, -2.5, -0.5, bs, R, 3),
                        _u(g, 0.5, 2.5, bs, R, 3), _u(g, 0, 1, bs, R, 3), _u(g, 0.3, 0.8, bs, R, 3)], -1)
    pre_stlp = stlp[:, :, :, None, :].contiguous()  # (bs,R,3,1,6)

    pw = _u(g, -0.5, 0.5, bs, R, 3, nt) * 0.1
    pa = _u(g, -5.0, 5.0, bs, R, 3, nt)
    params = torch.stack([pw, pa], -1)

    return {
        "ego_traj": ego_traj.contiguous(),
        "neighbors": neighbors.contiguous(),
        "neighbors_traj": ntraj.contiguous(),
        "currlane_wpts": lanes["curr"].contiguous(),
        "leftlane_wpts": lanes["left"].contiguous(),
        "rightlane_wpts": lanes["right"].contiguous(),
        "curr_id": ids["curr"], "left_id": ids["left"], "right_id": ids["right"],
        "gt_high_level": torch.randint(0, 3, (bs, 1), generator=g).float(),
        "pre_stlp": pre_stlp,
        "params": params.contiguous(),
        "params_init": params.clone(),
        "tj_scores_prior": _u(g, -1, 1, bs, R, 3),
        "traj_i": torch.arange(bs).float(),
        "ti": torch.zeros(bs),
    }


def make_dense_stl_input(n, nt=20, dt=0.5, n_neighbors=8, n_segs=15, seed=1008, endcaps=False, overlap=False):
    """Config-1/5 shape: ``n`` pre-rolled trajectories with per-row (dense) scene tensors,
    exactly what the reference's compute_stl_dense consumes (nusc_train.py:318-345).
    ``endcaps``: a third of the rows start 20 m behind the first lane point and a third 85 m further along
    the lane (so they run off its far end) — the cases the --inline end-cap distances exist for.

    ``overlap``: every fifth row gets neighbour 0 riding 1.2 m ahead / 0.6 m beside the ego (overlapping boxes: the
    negative clearances and the --collision_loss term are exercised).

    Returns (stl_input dict, stl_idx (n,1), mask (n,)).
    """
    g = torch.Generator().manual_seed(int(seed))
    bs = max(1, math.ceil(n / 192))
    b = make_scene_batch(bs, nt=nt, dt=dt, n_neighbors=n_neighbors, n_segs=n_segs, n_randoms=64, seed=seed)
    m = 192
    rep = lambda x: x.unsqueeze(1).repeat((1, m) + (1,) * (x.dim() - 1)).reshape((-1,) + x.shape[1:])[:n]
    s0 = rep(b["ego_traj"][:, 0, :4])
    if endcaps:
        grp = (torch.arange(n) // 3) % 3
        shift = torch.where(grp == 1, torch.tensor(-20.0), torch.where(grp == 2, torch.tensor(85.0), torch.tensor(0.0)))
        s0 = s0.clone()
        s0[:, 0] = s0[:, 0] + shift * torch.cos(s0[:, 2])
        s0[:, 1] = s0[:, 1] + shift * torch.sin(s0[:, 2])
    w = _u(g, -0.05, 0.05, n, nt)
    a = _u(g, -5.0, 5.0, n, nt) * 0.2
    st = s0
    tr = [st]
    for t in range(nt - 1):
        x, y, th, v = tr[-1].unbind(-1)
        tr.append(torch.stack([x + v * torch.cos(th) * dt, y + v * torch.sin(th) * dt,
                               th + w[:, t] * dt, v + a[:, t] * dt], -1))
    ego = torch.stack(tr, 1).contiguous()
    stlp = b["pre_stlp"].reshape(bs, 64, 3, 6)[:, 0:1].repeat(1, 64, 1, 1).reshape(bs * m, 1, 6)[:n]
    valids = torch.cat([b["curr_id"], b["left_id"], b["right_id"]], -1)  # (bs,3)
    mask = valids[:, None, :].repeat(1, 64, 1).reshape(-1)[:n]
    idx = torch.tensor([0.0, 1.0, 2.0]).repeat(bs * 64)[:n].reshape(n, 1)
    nei = rep(b["neighbors_traj"]).contiguous()
    if overlap:
        rows = torch.arange(n) % 5 == 0
        c, sn = torch.cos(ego[rows, :, 2]), torch.sin(ego[rows, :, 2])
        nei[rows, 0] = torch.stack([torch.ones_like(c), ego[rows, :, 0] + 1.2 * c - 0.6 * sn, ego[rows, :, 1] + 1.2 * sn + 0.6 * c,
                                    ego[rows, :, 2] + 0.05, ego[rows, :, 3], torch.full_like(c, 4.5), torch.full_like(c, 1.9)], -1)
    stl_input = {
        "ego_traj": ego,
        "neighbors": nei,
        "currlane_wpts": rep(b["currlane_wpts"]).contiguous(),
        "leftlane_wpts": rep(b["leftlane_wpts"]).contiguous(),
        "rightlane_wpts": rep(b["rightlane_wpts"]).contiguous(),
        "stlp": stlp.contiguous(),
        "dense_valids": mask.clone(),
    }
    return stl_input, idx, mask


def make_weights(seed=1007, nt=20, n_segs=15, hidden=256, rect_hidden=256, rect_extra_in=0):
    """Random-init weights with the reference ``Net`` state_dict keys/shapes:
    {ego,neighbor,lane}_encoder, policy_net (303->256->256->2nt), merge_net (2nt->32->32->2nt),
    rect_net (271->256->256->2nt).  U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear's default,
    but drawn from an explicit seeded generator so every consumer sees identical tensors."""
    g = torch.Generator().manual_seed(int(seed))
    sd = {}

    def mlp(name, dims):
        for li, (i, o) in zip((0, 2, 4), zip(dims[:-1], dims[1:])):
            k = 1.0 / math.sqrt(i)
            sd["%s.%d.weight" % (name, li)] = (torch.rand(o, i, generator=g) * 2 - 1) * k
            sd["%s.%d.bias" % (name, li)] = (torch.rand(o, generator=g) * 2 - 1) * k

    out = nt * 2
    mlp("ego_encoder", [6, hidden, hidden, 32])
    mlp("neighbor_encoder", [7, hidden, hidden, 32])
    mlp("lane_encoder", [n_segs * 3, hidden, hidden, 32])
    mlp("policy_net", [224 + out + 32 + 1 + 6, hidden, hidden, out])
    mlp("merge_net", [out, 32, 32, out])
    # rect_extra_in = 2*nt for --diverse_fuse_type cat (drawn last: the other tensors do not depend on it)
    mlp("rect_net", [224 + 1 + 6 + out + rect_extra_in, rect_hidden, rect_hidden, out])
    return sd


def noise_stream(seed, n, dim, count):
    """The injected-noise sequence of the deterministic test mode: ``count`` tensors (n,dim)
    drawn in order from one seeded CPU generator (x_T first, then z for i=steps-1..2)."""
    g = torch.Generator().manual_seed(int(seed))
    return [torch.randn(n, dim, generator=g) for _ in range(count)]
