"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (PyTorch fp32 on the host, same operation order as upstream) of the hot
path of mengyuest/pSTL-diffusion-policy: STL robustness algebra, unicycle rollout, lane /
neighbour predicates, the driving pSTL spec, the MLP denoiser + DDPM reverse loop (with
STL-gradient guidance), best-of-K selection and the RefineNet head.  Every function cites
the reference file:line it follows (paths relative to the reference checkout).

Parity pin: the reference ships no tests or golden vectors ("parity unpinned" upstream).
This file is pinned instead against outputs of the reference itself, imported in the build
container by tests/golden/make_golden.py; the resulting fixtures live in tests/golden/*.npz
and tests/test_oracle_golden.py re-checks the oracle against them on every run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import math

import torch

NEG_INF = -float("inf")

# --------------------------------------------------------------------------------------
# STL algebra.  A formula is a nested tuple:
#   ("ap", fn) ("not", f) ("and", l, r) ("or", l, r) ("imply", l, r) ("listand", [f...])
#   ("always", ts, te, f) ("eventually", ts, te, f) ("once", ts, te, f)
#   ("untimed_until", l, r) ("until", ts, te, l, r)
# --------------------------------------------------------------------------------------


def smax(x, tau, hard=False, dim=1):
    """stl_d_lib.py:6-13 — logsumexp(x*tau)/tau along dim (keepdim); empty -> -inf."""
    if x.shape[1] == 0:
        return torch.full((x.shape[0], 1), NEG_INF, dtype=x.dtype)
    if hard:
        return x.max(dim=dim, keepdim=True)[0]
    return torch.logsumexp(x * tau, dim=dim, keepdim=True) / tau


def smin(x, tau, hard=False, dim=1):
    """stl_d_lib.py:15-19 — note the empty window is -inf here too."""
    if x.shape[1] == 0:
        return torch.full((x.shape[0], 1), NEG_INF, dtype=x.dtype)
    return -smax(-x, tau, hard, dim)


def smax2(a, b, tau, hard=False):
    """stl_d_lib.py:21-23."""
    return smax(torch.stack([a, b], dim=1), tau, hard).squeeze(1)


def smin2(a, b, tau, hard=False):
    """stl_d_lib.py:25-26."""
    return -smax2(-a, -b, tau, hard)


def _lim(v, T):
    return max(min(v, T), 0)


def stl_eval(f, x, tau, hard=False):
    """Robustness trace (N,T) of formula tuple ``f`` on input ``x`` (stl_d_lib.py:70-203)."""
    op = f[0]
    if op == "ap":
        return f[1](x)
    if op == "not":
        return -stl_eval(f[1], x, tau, hard)
    if op == "and":
        return smin2(stl_eval(f[1], x, tau, hard), stl_eval(f[2], x, tau, hard), tau, hard)
    if op == "or":
        return smax2(stl_eval(f[1], x, tau, hard), stl_eval(f[2], x, tau, hard), tau, hard)
    if op == "imply":  # stl_d_lib.py:136  Or(Not(l), r)
        return smax2(-stl_eval(f[1], x, tau, hard), stl_eval(f[2], x, tau, hard), tau, hard)
    if op == "listand":  # stl_d_lib.py:101-112
        v = torch.stack([stl_eval(c, x, tau, hard) for c in f[1]], dim=1)
        return smin(v, tau, hard)[:, 0]
    if op in ("always", "eventually", "once"):  # stl_d_lib.py:144-180, window [t+ts, t+te) clipped
        ts, te, s = f[1], f[2], stl_eval(f[3], x, tau, hard)
        T = s.shape[1]
        red = smin if op == "always" else smax
        return torch.cat([red(s[:, _lim(t + ts, T):_lim(t + te, T)], tau, hard) for t in range(T)], dim=-1)
    if op == "untimed_until":  # stl_d_lib.py:186-192 (hard flag only reaches the pair-min)
        ls, rs = stl_eval(f[1], x, tau, hard), stl_eval(f[2], x, tau, hard)
        inf_ls = -torch.logcumsumexp(-ls * tau, dim=1) / tau
        m = smin2(rs, inf_ls, tau, hard)
        return (torch.logcumsumexp(m.flip(1) * tau, dim=1) / tau).flip(1)
    if op == "until":  # stl_d_lib.py:194-203
        ts, te, l, r = f[1], f[2], f[3], f[4]
        if ts == 0:
            return stl_eval(("untimed_until", l, r), x, tau, hard)
        return stl_eval(("and", ("eventually", ts, te, r), ("always", 0, ts, ("untimed_until", l, r))), x, tau, hard)
    raise ValueError(op)


# --------------------------------------------------------------------------------------
# Rollout and predicates
# --------------------------------------------------------------------------------------


def rollout(s, us, dt):
    """nusc_train.py:29-49 — Euler unicycle; (...,4) x (...,T,2) -> (...,T+1,4)."""
    out = [s]
    for t in range(us.shape[-2]):
        c = out[-1]
        th, v = c[..., 2], c[..., 3]
        ds = torch.stack([v * torch.cos(th), v * torch.sin(th), us[..., t, 0], us[..., t, 1]], dim=-1)
        out.append(c + ds * dt)
    return torch.stack(out, dim=-2)


def lane_distance(points, lane, clip=False, inline=False):
    """nusc_api.py:685-739 (efficient branch, with_angle=True).
    points (N,T,3)=[x,y,th]; lane (N,nseg,3) -> signed distance (N,T), 1-cos heading error (N,T)."""
    n, nseg, _ = lane.shape
    t = points.shape[1]
    pd = torch.norm(points[..., None, :2] - lane[:, None, :, :2], dim=-1)
    mi = torch.argmin(pd[:, :, :-1] + pd[:, :, 1:], dim=2)
    gi = mi.unsqueeze(-1).repeat(1, 1, 3)
    p2 = torch.gather(lane, 1, gi)
    p3 = torch.gather(lane, 1, gi + 1)
    x1, y1 = points[..., 0], points[..., 1]
    x2, y2, x3, y3 = p2[..., 0], p2[..., 1], p3[..., 0], p3[..., 1]
    area = x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2)
    base = torch.norm((p2 - p3)[..., :2], dim=-1)
    l2 = torch.clamp((x1 - x2) ** 2 + (y1 - y2) ** 2, 1e-3) ** 0.5
    ok = (base != 0).float()
    d0 = ok * area / torch.clip(base, 1e-7) + (1 - ok) * l2
    if inline:  # nusc_api.py:716-724
        l2b = torch.clamp((x1 - x3) ** 2 + (y1 - y3) ** 2, 1e-3) ** 0.5
        behind = ((x1 - x2) * (x3 - x2) + (y1 - y2) * (y3 - y2) <= 0) & (mi == 0)
        ahead = ((x1 - x3) * (x2 - x3) + (y1 - y3) * (y2 - y3) <= 0) & (mi == nseg - 2)
        normal = ~(behind | ahead)
        d0 = normal * d0 + behind * l2 * torch.sign(d0) + ahead * l2b * torch.sign(d0)
    ang = 1 - torch.cos(p2[..., 2] - points[..., 2])
    if clip:
        d0 = torch.clip(d0, -5, 5)
    return d0.reshape(n, t), ang.reshape(n, t)


def _anchor_circles(x, y, th, L, W, nL=4, nW=1):
    """utils.py:465-497: the nL x nW grid of circle centres (row-major over (length, width)) and their radius."""
    r = torch.minimum(torch.maximum(L / nL / 2, W / nW / 2), W / 2)
    alpha = torch.linspace(0, 1, nL)
    beta = torch.linspace(0, 1, nW)  # num_W=1: beta=[0] picks the y3+r end
    xs_ = (-L / 2 + r)[..., None] * (1 - alpha) + (L / 2 - r)[..., None] * alpha
    ys_ = (-W / 2 + r)[..., None] * (1 - beta) + (W / 2 - r)[..., None] * beta
    lead = list(x.shape)
    xs_ = xs_[..., None].expand(lead + [nL, nW]).reshape(lead + [nL * nW])
    ys_ = ys_[..., None, :].expand(lead + [nL, nW]).reshape(lead + [nL * nW])
    cx = xs_ * torch.cos(th[..., None]) - ys_ * torch.sin(th[..., None]) + x[..., None]
    cy = xs_ * torch.sin(th[..., None]) + ys_ * torch.cos(th[..., None]) + y[..., None]
    return torch.stack([cx, cy], dim=-1), r


def neighbour_clearance(ego, nei, ego_L=4.084, ego_W=1.730, nL=4, nW=1, full=False):
    """utils.py:499-526 + nusc_train.py:142-148.
    ego (N,T,>=3); nei (N,K,T,7)=[valid,x,y,th,v,L,W] -> min_k clearance (N,T); ``full`` (--collision_loss,
    nusc_train.py:81-83) also returns min_centroid_d (N,K,T) and radius_sum (N,K,T)."""
    e = ego.unsqueeze(1)
    c1, r1 = _anchor_circles(e[..., 0], e[..., 1], e[..., 2], ego_L * torch.ones_like(e[..., 0]),
                             ego_W * torch.ones_like(e[..., 0]), nL, nW)
    c2, r2 = _anchor_circles(nei[..., 1], nei[..., 2], nei[..., 3], nei[..., 5], nei[..., 6], nL, nW)
    d = torch.norm(c1[..., None, :] - c2[..., None, :, :], dim=-1)
    md = d.reshape(list(d.shape[:-2]) + [(nL * nW) ** 2]).min(dim=-1)[0]
    ind = nei[..., 0]
    clear = torch.min(torch.clip(md - r1 - r2, -5, 20) * ind + (1 - ind) * 100, dim=1)[0]
    if full:
        return clear, md * ind + (1 - ind) * 100, r1 + r2
    return clear


def predicates(x, ego_L=4.084, ego_W=1.730, clip_dist=False, inline=False, nL=4, nW=1, collision=False):
    """nusc_train.py:74-93 — adds the lane and neighbour signals to the dense dict ``x``."""
    p = x["ego_traj"][..., 0:3]
    for k in ("curr", "left", "right"):
        x["x2%s_d" % k], x["x2%s_th" % k] = lane_distance(p, x["%slane_wpts" % k], clip_dist, inline)
    if collision:
        x["min_nei_d"], x["min_centroid_d"], x["radius_sum"] = neighbour_clearance(x["ego_traj"], x["neighbors"], ego_L,
                                                                                   ego_W, nL, nW, full=True)
    else:
        x["min_nei_d"] = neighbour_clearance(x["ego_traj"], x["neighbors"], ego_L, ego_W, nL, nW)
    return x


def collision_loss(x, weight):
    """nusc_train.py:416-420 on the signals ``predicates(..., collision=True)`` added."""
    coll = torch.relu(1 - x["min_centroid_d"] / torch.clip(x["radius_sum"], 1e-1))
    return torch.mean(torch.clip(torch.sum(coll, dim=-1), max=1)) * weight


def guidance_steps(steps, before=1000, sets=None, freq=None, reverse=False):
    """nusc_train.py:589-598: the reverse steps i that run guidance."""
    hit = []
    for i in range(1, steps):
        i_val = steps - 1 - i if reverse else i
        if sets is not None:
            on = i_val in sets
        elif freq is not None:
            on = i_val % freq == 0
        else:
            on = i <= before
        if on:
            hit.append(i)
    return hit


def driving_spec(nt, norm=False):
    """nusc_train.py:95-140: [stl_curr, stl_left, stl_right] as formula tuples; ``norm`` = --norm_stl (:88-91,
    98-113: margins divided by clip(vmax-vmin,.3), clip(5(dmax-dmin),.3), clip(dsafe,.3))."""
    G = lambda f: ("always", 0, nt, f)
    one = lambda x: 1.0
    vf = (lambda x: torch.clip(x["stlp"][..., 1] - x["stlp"][..., 0], 0.3)) if norm else one
    df = (lambda x: torch.clip((x["stlp"][..., 3] - x["stlp"][..., 2]) * 5, 0.3)) if norm else one
    sf = (lambda x: torch.clip(x["stlp"][..., 4], 0.3)) if norm else one
    v_lo = G(("ap", lambda x: (x["ego_traj"][..., 3] - x["stlp"][..., 0]) / vf(x)))
    v_hi = G(("ap", lambda x: (-x["ego_traj"][..., 3] + x["stlp"][..., 1]) / vf(x)))
    d_lo = G(("ap", lambda x: (x["x2curr_d"] - x["stlp"][..., 2]) / df(x)))
    d_hi = G(("ap", lambda x: (-x["x2curr_d"] + x["stlp"][..., 3]) / df(x)))
    th_c = G(("ap", lambda x: (x["stlp"][..., 5] - x["x2curr_th"]) / x["stlp"][..., 5]))
    safe = G(("ap", lambda x: (x["min_nei_d"] - x["stlp"][..., 4]) / sf(x)))

    def reach(side):
        band = ("and", ("ap", lambda x: (x["x2%s_d" % side] - x["stlp"][..., 2]) / df(x)),
                ("ap", lambda x: (-x["x2%s_d" % side] + x["stlp"][..., 3]) / df(x)))
        rd = ("eventually", 0, nt // 2, G(band))
        rt = ("eventually", 0, nt // 2,
              G(("ap", lambda x: (x["stlp"][..., 5] - x["x2%s_th" % side]) / x["stlp"][..., 5])))
        return rd, rt

    ld, lt = reach("left")
    rd, rt = reach("right")
    return [("listand", [v_lo, v_hi, d_lo, d_hi, th_c, safe]),
            ("listand", [v_lo, v_hi, ld, lt, safe]),
            ("listand", [v_lo, v_hi, rd, rt, safe])]


def mask_mean(v, m):
    """nusc_train.py:23-27."""
    return torch.mean(v * m) / torch.clip(torch.mean(m), 1e-2)


def stl_scores(x, mode, tau=100.0, nt=None, norm=False, **pred_kw):
    """nusc_train.py:318-323,150-151 — all three formulas, [:,0], arithmetic select (+outlier 1.0).
    ``x`` dense dict; ``mode`` (N,) float in {0,1,2,3}.  Returns scores (N,)."""
    x = predicates(x, **pred_kw)
    T = x["ego_traj"].shape[1] if nt is None else nt
    per = [stl_eval(f, x, tau)[:, 0] for f in driving_spec(T, norm)]
    per.append(per[-1].detach() * 0.0 + 1.0)
    return sum(per[k] * (mode == k).float() for k in range(4))


# --------------------------------------------------------------------------------------
# Denoiser, sampler, RefineNet, selection.  ``W`` is a state_dict-like mapping with the
# reference key names ({ego,neighbor,lane}_encoder.*, policy_net.*, merge_net.*, rect_net.*).
# --------------------------------------------------------------------------------------


def mlp3(W, name, x):
    """utils.py:91-101 — Linear/ReLU/Linear/ReLU/Linear with keys name.{0,2,4}."""
    lin = torch.nn.functional.linear
    h = torch.relu(lin(x, W[name + ".0.weight"], W[name + ".0.bias"]))
    h = torch.relu(lin(h, W[name + ".2.weight"], W[name + ".2.bias"]))
    return lin(h, W[name + ".4.weight"], W[name + ".4.bias"])


def to_ego_frame(state, base, valid=None):
    """nusc_model.py:238-263."""
    bx, by, bth = base[..., 0], base[..., 1], base[..., 2]
    if valid is not None:
        xt, yt, tr = state[..., 0] - bx * valid, state[..., 1] - by * valid, state[..., 2] - bth * valid
    else:
        xt, yt, tr = state[..., 0] - bx, state[..., 1] - by, state[..., 2] - bth
    return torch.stack([xt * torch.cos(bth) + yt * torch.sin(bth), -xt * torch.sin(bth) + yt * torch.cos(bth), tr], -1)


def encode_scene(W, b):
    """nusc_model.py:55-95 — (bs,224) feature: ego32 | nei min/mean/max 96 | 3 lanes 96."""
    bs = b["ego_traj"].shape[0]
    ego = b["ego_traj"][:, 0]
    eu = ego.unsqueeze(1)
    nb = b["neighbors"]
    nin = torch.cat([nb[..., 0:1], to_ego_frame(nb[..., 1:4], eu, nb[..., 0]), nb[..., 4:7]], -1)
    ln = torch.stack([to_ego_frame(b["%slane_wpts" % k], eu, b["%s_id" % k]) for k in ("curr", "left", "right")], 1)
    lin_ = torch.cat([ln[..., 0:1, :], ln[..., 1:, :] - ln[..., :-1, :]], dim=-2).reshape(bs, 3, -1)
    ein = torch.cat([to_ego_frame(ego[..., :3], ego[..., :3]), ego[..., 3:]], -1)
    ef = mlp3(W, "ego_encoder", ein)
    nf = mlp3(W, "neighbor_encoder", nin)
    nf = torch.cat([nf.min(1)[0], nf.mean(1), nf.max(1)[0]], -1)
    lf = mlp3(W, "lane_encoder", lin_).reshape(bs, -1)
    return torch.cat([ef, nf, lf], -1)


def time_embedding(t, channels=32):
    """nusc_model.py:48-53; t (N,1) integer tensor."""
    inv = 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels))
    a = t.repeat(1, channels // 2) * inv
    return torch.cat([torch.sin(a), torch.cos(a)], -1)


def eps_model(W, feat_dense, x, t, hl, stlp):
    """nusc_model.py:118-162 — policy_net([feat,x,temb,hl,stlp]) + x."""
    inp = torch.cat([feat_dense, x, time_embedding(t), hl, stlp], -1)
    return mlp3(W, "policy_net", inp) + x


def ddpm_schedule(steps=100):
    """nusc_train.py:528-537 (cos is always forced on, :1782)."""
    t = torch.linspace(0, 1, steps + 1)
    ab = torch.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    beta = torch.clip(1 - ab[1:] / ab[:-1], 0, 0.999) * 0.2
    alpha = 1.0 - beta
    return beta, alpha, torch.cumprod(alpha, 0)


def to_controls(x, nt, w_max=0.5, a_max=5.0, clip=True):
    """nusc_train.py:647-655."""
    x = x.reshape(-1, nt, 2)
    w, a = x[..., 0] * w_max, x[..., 1] * a_max
    if clip:
        w, a = torch.clip(w, -w_max, w_max), torch.clip(a, -a_max, a_max)
    return torch.stack([w, a], -1)


def guidance_update(mu, beta_t, dense, s0, mode, valid, lr, thres, nt, dt, tau=100.0, niters=1,
                    w_max=0.5, a_max=5.0):
    """nusc_train.py:599-627 — fresh Adam per reverse step, ``niters`` steps.

    Upstream quirk reproduced on purpose: ``mu_opt = mu_init.detach().requires_grad_()`` (:606)
    SHARES STORAGE with ``mu_init``, so the in-place Adam step also moves ``mu_init``; the
    "clip |delta| to beta_t" line (:625-626) therefore sees delta == 0 on the first iteration
    and the net effect of niters=1 is the plain first Adam step  mu -= lr*g/(|g|+1e-8).
    From the second iteration on ``mu_opt.data`` is a fresh tensor and the clip acts relative
    to the once-stepped ``mu_init``."""
    N = s0.shape[0]
    mu_init = mu.reshape(N, nt, 2)
    mu_opt = mu_init.detach().requires_grad_()  # aliases mu_init, exactly as upstream
    opt = torch.optim.Adam([mu_opt], lr=lr)
    for _ in range(niters):
        u = torch.stack([mu_opt[..., 0] * w_max, mu_opt[..., 1] * a_max], -1)
        tr = rollout(s0, u, dt)
        x = dict(dense)
        x["ego_traj"] = tr[:, :-1]
        sc = stl_scores(x, mode, tau)
        loss = mask_mean(torch.relu(thres - sc), valid.reshape(-1))
        opt.zero_grad()
        loss.backward()
        opt.step()
        with torch.no_grad():
            mu_opt.data = mu_init + torch.clip(torch.abs(mu_opt - mu_init), -beta_t, beta_t)
    return mu_opt.reshape(N, -1).detach()


def ddpm_sample(W, feat_dense, hl, stlp, x_T, noises, steps=100, nt=20, clip=True, guidance=None):
    """nusc_train.py:557-645 — reverse loop i=steps-1..1 with t==i.
    ``noises``: list of (N,2nt) tensors consumed in order for i>1 (injected z).
    ``guidance``: None or dict(before | sets | freq [, reverse], lr, thres, dense, s0, mode, valid, dt, tau, niters).
    Returns list of steps (N,nt,2) control iterates x_T..x_0 (normalised as diff_full does)."""
    beta, alpha, abar = ddpm_schedule(steps)
    n = x_T.shape[0]
    x = x_T
    its = [x]
    zi = 0
    guided = set() if guidance is None else set(guidance_steps(steps, guidance.get("before", 1000), guidance.get("sets"),
                                                               guidance.get("freq"), guidance.get("reverse", False)))
    for i in reversed(range(1, steps)):
        t = torch.full((n, 1), i, dtype=torch.long)
        with torch.no_grad():
            eps = eps_model(W, feat_dense, x, t, hl, stlp)
        a, ah, b = alpha[i], abar[i], beta[i]
        if i > 1:
            z = noises[zi]
            zi += 1
        else:
            z = torch.zeros_like(x)
        mu = 1 / torch.sqrt(a) * (x - ((1 - a) / (torch.sqrt(1 - ah))) * eps)
        if i in guided:
            g = guidance
            mu = guidance_update(mu, b.item(), g["dense"], g["s0"], g["mode"], g["valid"], g["lr"], g["thres"],
                                 nt, g["dt"], g.get("tau", 100.0), g.get("niters", 1))
        x = mu + torch.sqrt(b) * z
        its.append(x)
    return [to_controls(r, nt, clip=clip) for r in its]


def refine(W, feat_dense, hl, stlp, u0, scores, n_randoms=64, n_shards=4, nt=20, w_max=0.5, a_max=5.0):
    """nusc_model.py:182-235 (diverse_loss, fuse=add, interval, no clip_rect)."""
    n = feat_dense.shape[0]
    g = mlp3(W, "merge_net", u0.reshape(-1, nt * 2))
    bs = int(u0.shape[0] / 3 / n_randoms)
    g = g.reshape(bs, n_randoms, 3, nt * 2).permute(0, 2, 1, 3)
    per = n_randoms // n_shards
    g = g.reshape(bs, 3, n_shards, per, nt * 2).max(dim=3, keepdim=True)[0]
    g = g.repeat(1, 1, 1, per, 1).reshape(bs, 3, n_randoms, nt * 2).permute(0, 2, 1, 3).reshape(u0.shape[0], nt, 2)
    fused = u0 + g
    raw = mlp3(W, "rect_net", torch.cat([feat_dense, hl, stlp, fused.reshape(n, nt * 2)], -1)).reshape(n, nt, 2)
    r = torch.tanh(raw)
    iw, ia = u0[..., 0], u0[..., 1]
    wm, am = (r[..., 0] >= 0).float(), (r[..., 1] >= 0).float()
    w = r[..., 0] * (iw + w_max) * (1 - wm) + r[..., 0] * (w_max - iw) * wm
    a = r[..., 1] * (ia + a_max) * (1 - am) + r[..., 1] * (a_max - ia) * am
    return u0 + torch.stack([w, a], -1) * (scores < 0).float()[:, None, None]


def densify(b, S, nt):
    """nusc_train.py:724-754 + :975-976 for the --load_stlp sampling path: one row per chain,
    flat index n=(scene*S+sample)*3+mode."""
    bs = b["currlane_wpts"].shape[0]
    m = S * 3
    dup = lambda x: x.unsqueeze(1).repeat((1, m) + (1,) * (x.dim() - 1)).reshape((-1,) + x.shape[1:])
    R = b["pre_stlp"].shape[1]
    stlp = b["pre_stlp"].reshape(bs, R, 3, 6)[:, 0:1].repeat(1, S, 1, 1).reshape(bs * m, 1, 6)
    valids = torch.cat([b["curr_id"], b["left_id"], b["right_id"]], -1)
    return {
        "neighbors": dup(b["neighbors_traj"][..., :7]),
        "currlane_wpts": dup(b["currlane_wpts"]),
        "leftlane_wpts": dup(b["leftlane_wpts"]),
        "rightlane_wpts": dup(b["rightlane_wpts"]),
        "stlp": stlp,
        "dense_valids": valids.unsqueeze(1).repeat(1, S, 1).reshape(bs * S, 3),
        "mode": torch.tensor([0.0, 1.0, 2.0]).repeat(bs * S),
        "s0": dup(b["ego_traj"][:, 0, :4]),
    }


def score_controls(dense, u, dt, tau=100.0):
    """rollout + STL score of controls u (N,nt,2) on dense scene rows (nusc_train.py:1101-1103)."""
    tr = rollout(dense["s0"], u, dt)
    x = {k: dense[k] for k in ("neighbors", "currlane_wpts", "leftlane_wpts", "rightlane_wpts", "stlp")}
    x["ego_traj"] = tr[:, :-1]
    return stl_scores(x, dense["mode"], tau), tr


REFINEMENT_ITERATES = {8: [0, 50, 80, 85, 90, 95, 98]}  # nusc_train.py:1051-1054, K = 8


def refinement(d, u, its, dt, tau=100.0, K=8, n_iters=50, thres=0.0005, lr=3e-1):
    """--refinement, nusc_train.py:1034-1071: violating rows become a softmax-weighted mix of their controls and K-1
    earlier iterates; the logits take 50 Adam steps on mask_mean(relu(5e-4 - score), valid)."""
    N = u.shape[0]
    lam = torch.ones(N, K, requires_grad=True)
    opt = torch.optim.Adam([lam], lr=lr)
    sc, _ = score_controls(d, u, dt, tau)
    valid = d["dense_valids"].reshape(-1)
    viol = ((sc <= 0) & (valid > 0)).float().reshape(N, 1, 1)
    for _ in range(n_iters):
        r = torch.softmax(lam, dim=-1)
        comb = [its[idx].detach() * r[..., j + 1:j + 2, None] for j, idx in enumerate(REFINEMENT_ITERATES[K])]
        oc = u.detach() * r[..., 0:1, None] + torch.sum(torch.stack(comb, dim=-1), dim=-1)
        oc = u.detach() * (1 - viol) + viol * oc
        s2, _ = score_controls(d, oc, dt, tau)
        loss = mask_mean(torch.relu(thres - s2), valid)
        opt.zero_grad()
        loss.backward()
        opt.step()
    return oc.detach()


def pipeline(W, b, x_T, noises, S=64, K=5, n_rolls=0, refinenet=True, steps=100, nt=20, dt=0.5, tau=100.0,
             guidance=None, n_randoms=64, n_shards=4, refine_mix=False):
    """The timed region of run_sampling_test (nusc_train.py:957-1105) for the README
    "Ours" / "Ours+guidance" flag sets.  Returns dict of intermediates for parity checks."""
    bs = b["ego_traj"].shape[0]
    d = densify(b, S, nt)
    feat = encode_scene(W, b)
    fd = feat.reshape(bs, 1, -1).repeat(1, S * 3, 1).reshape(bs * S * 3, -1)
    hl = d["mode"][:, None]
    stlp = d["stlp"][:, 0]
    gd = None
    if guidance is not None:
        gd = dict(guidance)
        gd.update(dense={k: d[k] for k in ("neighbors", "currlane_wpts", "leftlane_wpts", "rightlane_wpts", "stlp")},
                  s0=d["s0"], mode=d["mode"], valid=d["dense_valids"], dt=dt, tau=tau)
    its = ddpm_sample(W, fd, hl, stlp, x_T, noises, steps, nt, True, gd)
    # best-of-K over the last K iterates (nusc_train.py:992-1013)
    cand = torch.stack(its[-K:], 0)  # (K,N,nt,2)
    cs = torch.stack([score_controls(d, cand[k], dt, tau)[0] for k in range(K)], 0)
    best, bi = torch.max(cs, dim=0)
    u = cand[bi, torch.arange(bi.shape[0])]
    out = {"feature": feat, "final_iterate": its[-1], "cand_scores": cs, "best_idx": bi, "best_controls": u,
           "best_scores": best}
    if refinenet:
        with torch.no_grad():
            u = refine(W, fd, hl, stlp, u, best, n_randoms, n_shards, nt)
            for _ in range(n_rolls):
                sc, _ = score_controls(d, u, dt, tau)
                u = refine(W, fd, hl, stlp, u, sc, n_randoms, n_shards, nt)
    out["rect_controls"] = u
    if refine_mix:
        u = refinement(d, u, its, dt, tau)
    with torch.no_grad():
        sc, tr = score_controls(d, u, dt, tau)
    m = d["dense_valids"].reshape(-1)
    out.update(controls=u, scores=sc, trajs=tr, acc=mask_mean((sc > 0).float(), m))
    return out


def trajopt(b, S, nt, dt, iters, lr=0.005, thres=0.01, reg=10.0, w_max=0.5, a_max=5.0, tau=100.0, record=None):
    """Trajectory optimisation of the stored control parameters (nusc_train.py:287-316, 1303-1325):
    ``iters`` Adam steps on  mean(relu(thres - score) * valid) / clip(mean(valid), 1e-3)
    + reg * (mean(relu(w^2 - w_max^2)) + mean(relu(a^2 - a_max^2))),  controls ``b["params"]`` (bs,S,3,nt,2) used as is
    (no scaling, no clipping), pSTL parameters ``pre_stlp`` per chain (training layout, :742).
    Returns (params (N,nt,2), scores of the last evaluated iterate (N,)).  ``record(ii, loss, dense_loss, reg_loss,
    scores, grad_or_None, params_after)`` is called every iteration."""
    bs = b["currlane_wpts"].shape[0]
    m = S * 3
    N = bs * m
    dup = lambda x: x.unsqueeze(1).repeat((1, m) + (1,) * (x.dim() - 1)).reshape((-1,) + x.shape[1:])
    dense = {"neighbors": dup(b["neighbors_traj"][..., :7]), "currlane_wpts": dup(b["currlane_wpts"]),
             "leftlane_wpts": dup(b["leftlane_wpts"]), "rightlane_wpts": dup(b["rightlane_wpts"]),
             "stlp": b["pre_stlp"].reshape(N, 1, 6)}
    valid = torch.cat([b["curr_id"], b["left_id"], b["right_id"]], -1).unsqueeze(1).repeat(1, S, 1).reshape(bs * S, 3)
    s0 = dup(b["ego_traj"][:, 0, :4])
    p = b["params"].reshape(N, nt, 2).clone().requires_grad_()
    opt = torch.optim.Adam([p], lr=lr)
    spec = driving_spec(nt)
    scores = None
    for ii in range(iters):
        tr = rollout(s0, p, dt)
        x = dict(dense)
        x["ego_traj"] = tr[:, :-1]
        x = predicates(x)
        # formula i is evaluated on every row and column i of the (bs*S, 3) view is kept (:293-295)
        dense_scores = torch.stack([stl_eval(f, x, tau)[:, 0].reshape(bs * S, 3)[:, i] for i, f in enumerate(spec)], -1)
        dense_loss = torch.mean(torch.relu(thres - dense_scores) * valid) / torch.clip(torch.mean(valid), 1e-3)
        reg_loss = (torch.mean(torch.relu(p[..., 0] ** 2 - w_max ** 2)) + torch.mean(torch.relu(p[..., 1] ** 2 - a_max ** 2))) * reg
        loss = dense_loss + reg_loss
        opt.zero_grad()
        loss.backward()
        g = p.grad.detach().clone()
        opt.step()
        scores = dense_scores.detach().reshape(-1)
        if record is not None:
            record(ii, loss.item(), dense_loss.item(), reg_loss.item(), scores, g, p.detach().clone())
    return p.detach(), scores


def diversity(trajs, scores, valids, nt):
    """measure_diversity (nusc_api.py:817-877) restated with plain loops: per (scene, lane) the mean over the 2*nt
    way-point features of the population std over the samples with score > 0, and the summed scipy ConvexHull area of
    those samples' positions over the steps (0 on any Qhull error, for invalid lanes and when nothing is accepted).
    Returns (std (bs,3) float32, vol (bs,3) float64, ma_std_avg, ma_vol_avg)."""
    import numpy as np
    from scipy.spatial import ConvexHull
    tr = trajs.detach().cpu().numpy().astype(np.float32)
    acc = (scores.detach().cpu().numpy() > 0)
    val = valids.detach().cpu().numpy()[:, 0, :] != 0
    bs, m = tr.shape[0], tr.shape[1]
    std = np.zeros((bs, 3), np.float32)
    vol = np.zeros((bs, 3), np.float64)
    for b in range(bs):
        for l in range(3):
            sel = tr[b, acc[b, :, l], l]            # (n_acc, 2*nt)
            if sel.shape[0] > 0:
                std[b, l] = np.mean(np.std(sel, axis=0))
            if val[b, l] and sel.shape[0] > 0:
                for t in range(nt):
                    try:
                        vol[b, l] += ConvexHull(sel[:, 2 * t:2 * t + 2]).volume
                    except Exception:
                        pass
    n = max(int(val.sum()), 1)
    return std, vol, float((std * val).sum() / n), float((vol * val).sum() / n)



def refine_losses(rect, nn, scores, valid, n_scenes, S, nt, n_shards=4, diverse_loss=True, diverse_detach=False,
                  w_max=0.5, a_max=5.0, stl_nn_thres=0.0005, stl_weight=1.0, diversity_scale=1.0,
                  diversity_weight=1.0, rect_reg_loss=0.0, extra_rect_reg=0.0):
    """The RefineNet training losses of compute_policy_loss (nusc_train.py:411 loss_stl; :439-466 --diverse_loss:
    DPP diversity over groups of S/n_shards samples + masked regulariser; :468-478 the plain branch), as differentiable
    torch expressions of rect (N,nt,2) and scores (N,), rows n = (scene*S + sample)*3 + mode.  nn is a constant.
    Returns a dict with loss, loss_stl, loss_reg, loss_diversity, extra_loss_reg."""
    N = n_scenes * S * 3
    rect = rect.reshape(N, nt, 2)
    nn = nn.reshape(N, nt, 2).detach()
    lim = torch.tensor([w_max, a_max], dtype=rect.dtype)
    out = {}
    out["loss_stl"] = mask_mean(torch.relu(stl_nn_thres - scores), valid) * stl_weight
    zero = out["loss_stl"] * 0
    if diverse_loss:
        G = S // n_shards
        # (scene, sample, mode) -> (scene, mode, shard) groups of G consecutive samples
        u = (rect / lim).reshape(n_scenes, S, 3, nt * 2).transpose(1, 2).reshape(n_scenes * 3 * n_shards, G, nt * 2)
        qual = scores.reshape(n_scenes, S, 3).transpose(1, 2).reshape(n_scenes * 3 * n_shards, G)
        gap = (u.unsqueeze(2) - u.unsqueeze(1)).norm(dim=-1)
        kern = torch.exp(-diversity_scale * gap)
        pos = (qual > 0).to(rect.dtype)
        q = pos.detach() if diverse_detach else torch.exp(qual) * pos
        L = q.unsqueeze(2) * kern * q.unsqueeze(1)
        eye = torch.eye(G, dtype=rect.dtype).unsqueeze(0)
        div = (eye - torch.linalg.inv(L + eye)).diagonal(dim1=1, dim2=2).sum(1)
        out["loss_diversity"] = -div.mean() * diversity_weight
        keep = (scores.reshape(N, 1, 1) >= 0).to(rect.dtype)
        out["loss_reg"] = ((rect - nn) ** 2 * keep).mean() / torch.clip(keep.mean(), 1e-2)
        out["extra_loss_reg"] = zero
        out["loss"] = out["loss_stl"] + out["loss_reg"] * rect_reg_loss + out["loss_diversity"]
    else:
        d = (rect - nn) / lim
        out["loss_reg"] = ((d[..., 0] ** 2).mean() + (d[..., 1] ** 2).mean()) * rect_reg_loss
        v = (rect / lim) ** 2 - 1
        out["extra_loss_reg"] = (torch.relu(v[..., 0]).mean() + torch.relu(v[..., 1]).mean()) * extra_rect_reg
        out["loss_diversity"] = zero
        out["loss"] = out["loss_stl"] + out["loss_reg"] + out["extra_loss_reg"]
    return out


def refine_train_step(W, b, feat_scene, nn_controls, dt, tau=100.0, **loss_kw):
    """One --rect_head training step up to the gradients (nusc_train.py:1402-1405 rect_forward on the detached
    controls and their scores, :1420-1427 rollout + compute_policy_loss, :1523-1525 backward; the optimiser holds
    rect_net only, :1228-1233), pSTL parameters per chain (training layout, :742).  W: state_dict-keyed tensors; the six
    rect_net tensors are made leaves.  Returns rect, grad_rect, losses, prev_scores and grads {key: tensor}."""
    S, nt = loss_kw["S"], loss_kw["nt"]
    bs = b["currlane_wpts"].shape[0]
    m = S * 3
    N = bs * m
    dense = densify(b, S, nt)
    dense["stlp"] = b["pre_stlp"].reshape(N, 1, 6)
    keys = ["rect_net.%d.%s" % (li, k) for li in (0, 2, 4) for k in ("weight", "bias")]
    W = dict(W)
    for k in keys:
        W[k] = W[k].detach().clone().requires_grad_()
    nn_controls = nn_controls.reshape(N, nt, 2).detach()
    with torch.no_grad():
        prev_scores, _ = score_controls(dense, nn_controls, dt, tau)
    feat = feat_scene.unsqueeze(1).repeat(1, m, 1).reshape(N, -1)
    rect = refine(W, feat, dense["mode"].reshape(N, 1), dense["stlp"].reshape(N, 6), nn_controls, prev_scores,
                  n_randoms=S, n_shards=loss_kw["n_shards"], nt=nt, w_max=loss_kw["w_max"], a_max=loss_kw["a_max"])
    rect.retain_grad()
    scores, _ = score_controls(dense, rect, dt, tau)
    out = refine_losses(rect, nn_controls, scores, dense["dense_valids"].reshape(-1), **loss_kw)
    out["loss"].backward()
    return {"rect": rect.detach(), "grad_rect": rect.grad, "losses": out, "prev_scores": prev_scores,
            "grads": {k: W[k].grad for k in keys}, "W": W}


def ddpm_train_step(W, b, noise, steps, noised, S, nt, stl_bc_mask=True):
    """One denoiser training step up to the gradients (README step 1): net(batch, timestep per row, noised commands)
    (nusc_train.py:1352-1356, nusc_model.py:97-162) and loss_diffusion = mask_mean((noise - eps)^2, tj_scores_prior *
    valids > 0) (:435-437; the plain mean without stl_bc_mask), every encoder /
    policy_net tensor a leaf.  noise / steps / noised are diffusion_prep's outputs (:539-555).  pSTL per chain from
    pre_stlp (training layout).  Returns eps, feature (bs,224), loss and grads {key: tensor}."""
    bs = b["currlane_wpts"].shape[0]
    m = S * 3
    N = bs * m
    keys = [k for k in W if k.startswith(("ego_encoder", "neighbor_encoder", "lane_encoder", "policy_net"))]
    W = dict(W)
    for k in keys:
        W[k] = W[k].detach().clone().requires_grad_()
    feat = encode_scene(W, b)
    feat_dense = feat.unsqueeze(1).repeat(1, m, 1).reshape(N, -1)
    mode = torch.tensor([0.0, 1.0, 2.0]).repeat(bs * S).reshape(N, 1)
    eps = eps_model(W, feat_dense, noised, steps.reshape(N, 1), mode, b["pre_stlp"].reshape(N, 6))
    if stl_bc_mask:  # nusc_train.py:435-437: dense_scores = tj_scores_prior (:1282-1283), mask_mean (:23-27)
        valids = torch.cat([b["curr_id"], b["left_id"], b["right_id"]], dim=-1).unsqueeze(1).repeat(1, S, 1).reshape(N, 1)
        mk = (b["tj_scores_prior"].reshape(N, 1) * valids > 0).float()
        loss = torch.mean((noise - eps) ** 2 * mk) / torch.clip(torch.mean(mk), 1e-2)
    else:
        loss = torch.mean((noise - eps) ** 2)
    loss.backward()
    return {"eps": eps.detach(), "feature": feat.detach(), "loss": loss.detach(), "grads": {k: W[k].grad for k in keys}, "W": W}
