"""CPU tier: the per-trajectory math the CUDA kernels run (csrc/*.cuh, compiled for the host by
tests/hostsim) against the golden fixtures produced from the unmodified reference."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import pstl_b200  # noqa: F401
from pstl_b200 import stl_d_lib as S, synthetic, native
from pstl_b200.nusc_train import build_stl_cache, default_args
from formulas import recipes
from make_golden import kat_inputs
from hostsim.loader import load


def fp(a):
    return a.ctypes.data_as(C.c_void_p)


def ops_array(ops):
    return (native.Op * len(ops))(*[native.Op(*o) for o in ops])


def close(a, b, rtol=1e-5, atol=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(b)
    assert (np.isfinite(a) == fin).all()
    assert (a[~fin] == b[~fin]).all()
    if atol is None:
        atol = rtol * max(1.0, float(np.abs(b[fin]).max()) if fin.any() else 1.0)
    np.testing.assert_allclose(a[fin], b[fin], rtol=rtol, atol=atol)


def test_interpreter_kats(golden_dir):
    hs = load()
    G = np.load(os.path.join(golden_dir, "stl_kats.npz"))
    x0 = kat_inputs()
    N, T = x0["a"].shape
    fs = recipes(S)
    for name in [str(n) for n in G["names"]]:
        ops, leaves = S.compile_formula(fs[name])
        sig = np.ascontiguousarray(torch.stack([l.expression(x0) for l in leaves], 1).numpy(), np.float32)
        P = len(leaves)
        for tau in (1.0, 100.0):
            for hard in (0, 1):
                key = "%s|%g|%d" % (name, tau, hard)
                out = np.zeros((N, T), np.float32)
                gtr = np.zeros((N, T), np.float32)
                gtr[:, 0] = 1.0
                gsig = np.zeros_like(sig)
                want_g = (key + "|ga") in G.files
                rc = hs.hs_stl_signals(ops_array(ops), len(ops), P, T, T, fp(sig), N, C.c_float(tau), hard, fp(out),
                                       fp(gtr) if want_g else None, fp(gsig) if want_g else None)
                assert rc == 0
                close(out, G[key])
                if want_g:
                    got = {k: np.zeros((N, T), np.float32) for k in "abc"}
                    for i, l in enumerate(leaves):
                        got[l.comment] += gsig[:, i]
                    for k in "abc":
                        close(got[k], G[key + "|g" + k], atol=2e-6)


def test_need_t_one_matches_full(golden_dir):
    """demand-driven evaluation (need_t=1) gives the same t=0 robustness as the full trace"""
    hs = load()
    x0 = kat_inputs()
    N, T = x0["a"].shape
    fs = recipes(S)
    for name in ("nested_mix", "ev_alw_and", "until_2_5", "alw_ev"):
        ops, leaves = S.compile_formula(fs[name])
        sig = np.ascontiguousarray(torch.stack([l.expression(x0) for l in leaves], 1).numpy(), np.float32)
        full = np.zeros((N, T), np.float32)
        one = np.zeros((N, 1), np.float32)
        assert hs.hs_stl_signals(ops_array(ops), len(ops), len(leaves), T, T, fp(sig), N, C.c_float(100.0), 0, fp(full), None, None) == 0
        assert hs.hs_stl_signals(ops_array(ops), len(ops), len(leaves), T, 1, fp(sig), N, C.c_float(100.0), 0, fp(one), None, None) == 0
        np.testing.assert_array_equal(full[:, 0], one[:, 0])


@pytest.mark.parametrize("tag,n,nt,knei,seed", [("t20k8", 192, 20, 8, 1008), ("t50k16", 48, 50, 16, 1009)])
def test_fused_dense_scores_and_grads(golden_dir, tag, n, nt, knei, seed):
    hs = load()
    G = np.load(os.path.join(golden_dir, "stl_dense.npz"))
    x, idx, mask = synthetic.make_dense_stl_input(n, nt=nt, n_neighbors=knei, seed=seed)
    args = default_args(nt=nt)
    stls = build_stl_cache(args)
    progs = [S.compile_formula(f, fused=True)[0] for f in stls]
    arrs = [ops_array(p) for p in progs]
    ops3 = (C.c_void_p * 3)(*[C.cast(a, C.c_void_p) for a in arrs])
    nops = (C.c_int * 3)(*[len(p) for p in progs])
    f = lambda t: np.ascontiguousarray(t.numpy(), np.float32)
    nei, l0, l1, l2 = f(x["neighbors"]), f(x["currlane_wpts"]), f(x["leftlane_wpts"]), f(x["rightlane_wpts"])
    ego, stlp, mode = f(x["ego_traj"]), f(x["stlp"][:, 0]), f(idx[:, 0])
    scores = np.zeros(n, np.float32)
    # guidance-style loss gradient: d loss/d score = -[thres-score>0]*mask/(n*clip(mean(mask),1e-2))
    sc_ref = G[tag + "|scores"]
    m = mask.numpy()
    gs = np.where(0.0005 - sc_ref > 0, -m / (n * max(m.mean(), 1e-2)), 0.0).astype(np.float32)
    gego = np.zeros((n, nt, 4), np.float32)
    rc = hs.hs_score(ops3, nops, nt, knei, 15, fp(nei), fp(l0), fp(l1), fp(l2), 1, fp(mode), None, None, fp(ego), 4,
                     fp(stlp), n, C.c_float(0.5), C.c_float(100.0), C.c_float(1.0), C.c_float(1.0), 0, 0, fp(gs),
                     fp(scores), None, fp(gego))
    assert rc == 0
    close(scores, sc_ref)
    # the streaming plan evaluator (score_stream.cuh) on the same rows
    s2 = np.zeros(n, np.float32)
    rc = hs.hs_score_stream(ops3, nops, nt, knei, 15, fp(nei), fp(l0), fp(l1), fp(l2), 1, fp(mode), None, None, fp(ego), 4,
                            fp(stlp), n, C.c_float(0.5), C.c_float(100.0), C.c_float(1.0), C.c_float(1.0), 0, fp(s2))
    assert rc == 0
    close(s2, sc_ref)
    gref = G[tag + "|grad_ego"]
    np.testing.assert_allclose(gego, gref, rtol=2e-4, atol=2e-4 * np.abs(gref).max())
    # streaming forward + reverse sweep
    s3, g3 = np.zeros(n, np.float32), np.zeros((n, nt, 4), np.float32)
    rc = hs.hs_score_stream_grad(ops3, nops, nt, knei, 15, fp(nei), fp(l0), fp(l1), fp(l2), 1, fp(mode), None, None, fp(ego),
                                 4, fp(stlp), n, C.c_float(0.5), C.c_float(100.0), C.c_float(1.0), C.c_float(1.0), 0, fp(gs),
                                 fp(s3), None, fp(g3))
    assert rc == 0
    close(s3, sc_ref)
    np.testing.assert_allclose(g3, gref, rtol=2e-4, atol=2e-4 * np.abs(gref).max())


def _typed_leaves(nt):
    """typed predicate leaves of the driving spec (nusc_train.build_stl_cache) for hand-built formulas"""
    from pstl_b200 import nusc_train as NT
    P = S.AP.predicate
    one = lambda x: 1.0
    return {
        "vmin": P(one, native.SIG_V, 0, NT.I_VMIN, 1),
        "vmax": P(one, native.SIG_V, 1, NT.I_VMAX, 0),
        "dmin_l": P(one, native.SIG_D_LEFT, 0, NT.I_DMIN, 1),
        "dmax_l": P(one, native.SIG_D_LEFT, 1, NT.I_DMAX, 0),
        "th_l": P(one, native.SIG_TH_LEFT, 1, NT.I_THMAX, 0, native.DEN_THMAX),
        "dmin_c": P(one, native.SIG_D_CURR, 0, NT.I_DMIN, 1),
        "safe": P(one, native.SIG_NEI, 0, NT.I_DSAFE, 1),
        "safe_n": P(one, native.SIG_NEI, 0, NT.I_DSAFE, 1, native.DEN_SFACTOR),
        "unsafe": P(one, native.SIG_NEI, 1, NT.I_DSAFE, 0),
    }


def _plan_info(hs, f, nt):
    ops = S.compile_formula(f, fused=True)[0]
    out = (C.c_int * 9)()
    assert hs.hs_plan_info(ops_array(ops), len(ops), nt, out) == 0
    return dict(zip(("valid", "n_terms", "listand", "lane", "n_tapes", "need_pose", "need_lane", "need_nei", "nei_term"), out))


def test_plan_recognition():
    """which programs the streaming scorer takes (stl_program.h: pstl_make_plan)"""
    hs = load()
    nt = 20
    infos = [_plan_info(hs, f, nt) for f in build_stl_cache(default_args(nt=nt))]
    assert [i["valid"] for i in infos] == [1, 1, 1]
    assert [i["lane"] for i in infos] == [0, 1, 2]
    assert [i["n_tapes"] for i in infos] == [0, 2, 2]
    assert [i["n_terms"] for i in infos] == [6, 5, 5]
    assert all(i["nei_term"] == 0 and i["need_pose"] == nt for i in infos)  # moved to slot 0 by pstl_make_plan
    L = _typed_leaves(nt)
    A, E = S.Always, S.Eventually
    assert _plan_info(hs, A(0, 7, L["vmin"]), nt) == dict(valid=1, n_terms=1, listand=0, lane=-1, n_tapes=0, need_pose=7,
                                                           need_lane=0, need_nei=0, nei_term=-1)
    # not of the closed form: inner window that is not a suffix, negation, two lanes in one program, Until
    assert _plan_info(hs, A(0, nt, E(0, 5, L["vmin"])), nt)["valid"] == 0
    assert _plan_info(hs, A(0, nt, S.Not(L["vmin"])), nt)["valid"] == 0
    assert _plan_info(hs, S.ListAnd([A(0, nt, L["dmin_l"]), A(0, nt, L["dmin_c"])]), nt)["valid"] == 0
    assert _plan_info(hs, S.Until(0, nt, L["vmin"], L["vmax"]), nt)["valid"] == 0
    # the value-aware neighbour bound only for a lone soft-min  nei - q  term
    assert _plan_info(hs, S.ListAnd([A(0, nt, L["safe"]), A(0, nt, L["vmin"])]), nt)["nei_term"] == 0
    assert _plan_info(hs, S.ListAnd([A(0, nt, L["unsafe"]), A(0, nt, L["vmin"])]), nt)["nei_term"] == -1
    assert _plan_info(hs, S.ListAnd([E(0, nt, L["safe"]), A(0, nt, L["vmin"])]), nt)["nei_term"] == -1
    assert _plan_info(hs, S.ListAnd([A(0, nt, L["safe"]), A(0, 5, L["safe_n"])]), nt)["nei_term"] == -1


def test_stream_plan_equals_interpreter_on_variant_specs():
    """streaming closed form == postfix interpreter for hand-built formulas of the plan shape
    (clipped / shifted windows, Or pairs, soft-max outer operators, single-term programs, normalised leaves)"""
    hs = load()
    n, nt, knei = 96, 20, 8
    x, idx, mask = synthetic.make_dense_stl_input(n, nt=nt, n_neighbors=knei, seed=77)
    L = _typed_leaves(nt)
    A, E = S.Always, S.Eventually
    variants = [
        [S.ListAnd([A(2, 15, L["vmin"]), E(0, 7, L["vmax"]), E(3, 9, A(0, nt, S.And(L["dmin_l"], L["dmax_l"]))),
                    A(0, nt, L["safe_n"])])] * 3,
        [S.ListAnd([E(0, nt // 2, E(0, nt, S.Or(L["dmin_l"], L["th_l"]))), A(-3, 30, L["th_l"]), A(0, nt, L["safe"])])] * 3,
        [A(0, nt, L["safe"]), E(1, 6, A(0, nt, L["th_l"])), S.ListAnd([A(5, 5, L["vmin"]), A(0, nt, L["vmax"])])],
        [S.ListAnd([A(0, nt, L["unsafe"]), A(0, 9, L["dmin_c"])]), E(0, nt, L["safe"]),
         S.ListAnd([A(0, nt, A(0, nt, L["safe"])), A(0, nt, L["vmin"])])],
    ]
    f = lambda t: np.ascontiguousarray(t.numpy(), np.float32)
    nei, l0, l1, l2 = f(x["neighbors"]), f(x["currlane_wpts"]), f(x["leftlane_wpts"]), f(x["rightlane_wpts"])
    ego, stlp, mode = f(x["ego_traj"]), f(x["stlp"][:, 0]), f(idx[:, 0])
    for stls in variants:
        progs = [S.compile_formula(g, fused=True)[0] for g in stls]
        arrs = [ops_array(p) for p in progs]
        ops3 = (C.c_void_p * 3)(*[C.cast(a, C.c_void_p) for a in arrs])
        nops = (C.c_int * 3)(*[len(p) for p in progs])
        ref, got = np.zeros(n, np.float32), np.zeros(n, np.float32)
        assert hs.hs_score(ops3, nops, nt, knei, 15, fp(nei), fp(l0), fp(l1), fp(l2), 1, fp(mode), None, None, fp(ego), 4,
                           fp(stlp), n, C.c_float(0.5), C.c_float(100.0), C.c_float(1.0), C.c_float(1.0), 0, 0, None,
                           fp(ref), None, None) == 0
        assert hs.hs_score_stream(ops3, nops, nt, knei, 15, fp(nei), fp(l0), fp(l1), fp(l2), 1, fp(mode), None, None,
                                  fp(ego), 4, fp(stlp), n, C.c_float(0.5), C.c_float(100.0), C.c_float(1.0),
                                  C.c_float(1.0), 0, fp(got)) == 0
        close(got, ref)
        # reverse mode through the rollout: d score / d controls (scaled + clipped controls, as the guidance call)
        g = torch.Generator().manual_seed(3)
        ctl = np.ascontiguousarray(((torch.rand(n, nt, 2, generator=g) * 2.4 - 1.2)).numpy(), np.float32)
        s0 = np.ascontiguousarray(ego[:, 0, :4])
        ones = np.ones(n, np.float32)
        sa, sb = np.zeros(n, np.float32), np.zeros(n, np.float32)
        ga, gb = np.zeros((n, nt, 2), np.float32), np.zeros((n, nt, 2), np.float32)
        assert hs.hs_score(ops3, nops, nt, knei, 15, fp(nei), fp(l0), fp(l1), fp(l2), 1, fp(mode), fp(s0), fp(ctl), None, 0,
                           fp(stlp), n, C.c_float(0.5), C.c_float(100.0), C.c_float(0.5), C.c_float(5.0), 1, 0, fp(ones),
                           fp(sa), fp(ga), None) == 0
        assert hs.hs_score_stream_grad(ops3, nops, nt, knei, 15, fp(nei), fp(l0), fp(l1), fp(l2), 1, fp(mode), fp(s0), fp(ctl),
                                       None, 0, fp(stlp), n, C.c_float(0.5), C.c_float(100.0), C.c_float(0.5), C.c_float(5.0),
                                       1, fp(ones), fp(sb), fp(gb), None) == 0
        close(sb, sa)
        fin = np.isfinite(sa)
        np.testing.assert_allclose(gb[fin], ga[fin], rtol=2e-4, atol=2e-4 * max(1e-6, np.abs(ga[fin]).max()))
