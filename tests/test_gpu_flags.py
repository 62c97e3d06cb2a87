"""GPU tier, part 2: the flags of SURVEY 8(b) beyond the README command lines, the sampler's other call modes and the
benchmarked bf16 path — each against fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py).

Tolerances (north_star: 1e-5 relative on the fp32 path, 2e-2 on the bf16 denoiser path):
  * ``close_elem``: ELEMENTWISE |a - b| <= 1e-5 * max(|b|, floor); floor is the scale below which "relative" loses
    meaning for that tensor (1.0 for robustness scores / clearances / lane distances in metres, stated at each call).
  * gradients: 2e-4 (the reference's own soft-max weights carry ~1e-4 relative rounding noise at tau*x ~ 2000).
"""
import os

import numpy as np
import pytest
import torch

import pstl_b200  # noqa: F401
from pstl_b200 import synthetic
from pstl_b200 import nusc_train as NT
from pstl_b200.nusc_model import Net
from test_gpu_parity import close, cuda

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def close_elem(a, b, rtol=1e-5, floor=1.0, what=""):
    a, b = np.asarray(npy(a), np.float64), np.asarray(npy(b), np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    fin = np.isfinite(b)
    assert (np.isfinite(a) == fin).all(), what
    err = np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), floor)
    assert err.size == 0 or err.max() <= rtol, "%s: max elementwise rel err %.3g (rtol %.1g, floor %g)" % (what, err.max(), rtol, floor)


# ------------------------------------------------------------------------------------------
# predicate flags: --inline, --clip_dist, --norm_stl, --refined_nL/nW, --collision_loss
# ------------------------------------------------------------------------------------------
VARIANTS = {"inline": dict(inline=True), "inline_clip": dict(inline=True, clip_dist=True), "norm": dict(norm_stl=True),
            "nl3w2": dict(refined_nL=3, refined_nW=2), "nl6": dict(refined_nL=6), "coll": dict(collision_loss=1.0)}


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_flag_variants_golden(golden_dir, tag):
    G = np.load(os.path.join(golden_dir, "flags.npz"))
    x, idx, mask = synthetic.make_dense_stl_input(96, nt=20, n_neighbors=8, seed=1010, endcaps=True, overlap=True)
    args = NT.default_args(nt=20, **VARIANTS[tag])
    stls = NT.build_stl_cache(args)
    xc = cuda(x)
    xc["ego_traj"] = xc["ego_traj"].clone().requires_grad_()
    _, scores, acc, xo = NT.compute_stl_dense(xc, stls, idx.cuda(), mask.cuda(), args, debug=True)
    # lane distances are differences of products of world coordinates (|x|,|y| up to ~150 m here): their fp32
    # cancellation noise is ~1e-5 * coordinate, so the scale of "relative" is the coordinate range, not the distance
    for k in ("x2curr_d", "x2left_d", "x2right_d"):
        close(xo[k], G[tag + "|" + k], what=k)
    close_elem(xo["min_nei_d"], G[tag + "|min_nei_d"], floor=1.0, what="min_nei_d")
    close_elem(scores, G[tag + "|scores"], rtol=2e-5, floor=1.0, what="scores")
    loss = NT.mask_mean(torch.relu(args.stl_nn_thres - scores), mask.cuda())
    if tag == "coll":
        close_elem(xo["min_centroid_d"], G["coll|min_centroid_d"], floor=1.0, what="min_centroid_d")
        close_elem(xo["radius_sum"], G["coll|radius_sum"], floor=1.0, what="radius_sum")
        coll_dist = torch.relu(1 - xo["min_centroid_d"] / torch.clip(xo["radius_sum"], 1e-1))
        coll = torch.mean(torch.clip(torch.sum(coll_dist, dim=-1), max=1)) * args.collision_loss
        np.testing.assert_allclose(float(coll.detach()), float(G["coll|loss_coll"]), rtol=1e-5)
        loss = loss + coll
    (g,) = torch.autograd.grad(loss, [xc["ego_traj"]])
    close(g[..., :4], G[tag + "|grad_ego"][..., :4], rtol=2e-4, what="grad_ego")


def test_inline_fused_equals_generic_path():
    """--inline through the fused scorer (typed leaves) and through prep_stl_cache + the generic formula kernels"""
    x, idx, mask = synthetic.make_dense_stl_input(96, nt=20, n_neighbors=8, seed=1010, endcaps=True)
    args = NT.default_args(nt=20, inline=True)
    stls = NT.build_stl_cache(args)
    _, fused, _ = NT.compute_stl_dense(cuda(x), stls, idx.cuda(), mask.cuda(), args)
    xd = NT.prep_stl_cache(cuda(x), args)
    per = [f(xd, args.smoothing_factor)[:, 0] for f in stls]
    per.append(per[-1] * 0 + 1)
    close_elem(fused, NT.get_stl_scores(per, idx.cuda()[:, 0]), rtol=2e-5, what="inline fused vs generic")


def test_nondefault_anchor_grid_refuses_fused_pipeline():
    args = NT.default_args(refined_nL=3)
    b = cuda(synthetic.make_scene_batch(2, seed=3))
    b["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
    pack = NT.augment_batch_data(NT.LazyBatch(b), None, args, n_randoms=64)["_pstl_pack"]
    with pytest.raises(NotImplementedError):
        NT.score_pack(pack, torch.zeros(pack.N, 20, 2, device="cuda"), args, None)


# ------------------------------------------------------------------------------------------
# a7 get_neighbor_trajs, a20 closed-loop pick
# ------------------------------------------------------------------------------------------
def test_get_neighbor_trajs_golden(golden_dir):
    G = np.load(os.path.join(golden_dir, "flags.npz"))
    b = synthetic.make_scene_batch(3, nt=20, n_randoms=4, seed=1011)
    nei = b["neighbors"].cuda()
    close_elem(NT.get_neighbor_trajs(nei, 20, 0.5), G["nei|short"], floor=1.0, what="short")
    close_elem(NT.get_neighbor_trajs(nei, 20, 0.5, full=True), G["nei|full"], floor=1.0, what="full")


def test_closed_loop_pick_golden(golden_dir):
    G = np.load(os.path.join(golden_dir, "flags.npz"))
    sc = torch.from_numpy(G["pick|scores"]).cuda()
    keep = sc.clone()
    idx, highest, ctrl, traj = NT.closed_loop_pick(sc, torch.from_numpy(G["pick|controls"]).cuda(),
                                                   torch.from_numpy(G["pick|trajs"]).cuda())
    assert int(idx) == int(G["pick|idx"]) and int(idx) % 3 == 0 and int(idx) == 30 * 1  # first of the two tied rows
    assert float(highest) == float(G["pick|highest"])
    assert np.array_equal(npy(ctrl), G["pick|ctrl"]) and np.array_equal(npy(traj), G["pick|traj"])
    assert torch.equal(sc, keep)  # the caller's scores are left alone


# ------------------------------------------------------------------------------------------
# sampler modes
# ------------------------------------------------------------------------------------------
def _pipeline(flags, seed, bs, precision="fp32", **over):
    S_, nt = 16, 20
    args = NT.default_args(flags, n_randoms=S_, sampling_size=S_, precision=precision, **over)
    batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S_, seed=seed)
    net = Net(args)
    net.load_state_dict(synthetic.make_weights(1007, nt=nt), strict=True)
    net = net.cuda()
    stream = synthetic.noise_stream(seed + 77, bs * S_ * 3, nt * 2, 99)
    args.inject_noise = [t.cuda() for t in stream]
    out = NT.sample_and_score(net, cuda(batch), NT.build_stl_cache(args), NT.get_diffusion_coeffs(args), args)
    return out, net, batch, args


def _guided_steps_teacher_forced(G, tag, out, args, beta):
    """Every guided reverse step of the reference run, in isolation: the reference's own mu going in, our gradient and
    updated mu against the reference's.  A row is EXEMPT from the 1e-5 bound only when its reference score sits within
    2e-5 of the relu threshold (relu'(thres - score) flips on rounding) — those rows are counted and bounded by lr."""
    pack, stls = out["pack"], NT.build_stl_cache(args)
    mu_in, g_ref, mu_out, sc = (G[tag + "|gstep_" + k] for k in ("mu_in", "grad", "mu_out", "scores"))
    steps = [i for i in range(99, 0, -1) if NT.guidance_step_mask(args)[i]]
    assert len(steps) == mu_in.shape[0]
    n_exempt = 0
    for j, i in enumerate(steps):
        mu = torch.from_numpy(mu_in[j]).cuda().contiguous()
        mu, grad, _ = NT.guidance_step(pack, mu, stls, args, float(beta[i]))
        edge = np.abs(args.stl_nn_thres - sc[j]) < 2e-5
        n_exempt += int(edge.sum())
        ok = ~edge
        gr, g0 = npy(grad).reshape(pack.N, -1), g_ref[j]
        # gradient: 5e-4 of the row's largest entry (soft-max weight rounding at tau*x ~ 2000; measured max 3.3e-4),
        # exactly zero on the rows where the reference's is
        scale = np.abs(g0).max(axis=1, keepdims=True)
        assert (np.abs(gr - g0)[ok] <= 5e-4 * scale[ok] + 1e-12).all(), (tag, i, "grad")
        # update: with niters = 1 the step is lr * g / (|g| + 1e-8); 1e-5 absolute on the normalised controls
        err = np.abs(npy(mu).reshape(pack.N, -1) - mu_out[j])
        assert err[ok].max() <= 1e-5, (tag, i, err[ok].max())
        assert err.max() <= 2 * args.guidance_lr + 1e-5
    return n_exempt, len(steps) * pack.N


def test_guided_steps_golden(golden_dir):
    """README "Ours+guidance": per-step teacher-forced parity of the guidance update (no percentile)"""
    G = np.load(os.path.join(golden_dir, "pipeline.npz"))
    out, net, batch, args = _pipeline(NT.GUIDANCE_FLAGS, 2002, 2)
    beta = NT.get_diffusion_coeffs(args)[0].cpu().numpy()
    n_exempt, n_rows = _guided_steps_teacher_forced(G, "guide", out, args, beta)
    assert n_exempt <= 0.02 * n_rows, (n_exempt, n_rows)
    # the whole run: each guided step reproduces the reference's update to < 1e-6 when it starts from the reference's mu
    # (above); started from our own iterate, a row that crosses relu' / an arg-min on a 1e-7 input difference moves by
    # up to lr per later guided step, so the run is held to 1e-5 on the median row and lr * (guided steps) on every row
    a, b = npy(out["final_iterate"]), G["guide|final_iterate"]
    err = (np.abs(a - b) / np.array([0.5, 5.0])).reshape(a.shape[0], -1).max(axis=1)
    assert np.median(err) < 1e-5 and err.max() < args.guidance_lr * 10, (np.median(err), err.max())


def test_guided_run_on_the_split_operand_engine(golden_dir):
    """README "Ours+guidance" with --precision f16x3: the unguided runs and the posterior means of the guided steps come from
    k_denoiser_tc3; the whole run is held to the bounds the fp32 path is held to above"""
    G = np.load(os.path.join(golden_dir, "pipeline.npz"))
    out, net, batch, args = _pipeline(NT.GUIDANCE_FLAGS, 2002, 2, precision="f16x3")
    a, b = npy(out["final_iterate"]), G["guide|final_iterate"]
    err = (np.abs(a - b) / np.array([0.5, 5.0])).reshape(a.shape[0], -1).max(axis=1)
    assert np.median(err) < 1e-5 and err.max() < args.guidance_lr * 10, (np.median(err), err.max())


@pytest.mark.parametrize("tag,seed,extra", [("freq", 2011, ["--guidance_freq", "7"]),
                                            ("sets", 2012, ["--guidance_sets", "3", "40", "41", "--guidance_reverse"])])
def test_guidance_triggers_golden(golden_dir, tag, seed, extra):
    G = np.load(os.path.join(golden_dir, "sampler_modes.npz"))
    out, net, batch, args = _pipeline(NT.GUIDANCE_FLAGS + extra, seed, 1)
    assert int(NT.guidance_step_mask(args).sum()) == int(G[tag + "|n_guidance_calls"])
    beta = NT.get_diffusion_coeffs(args)[0].cpu().numpy()
    n_exempt, n_rows = _guided_steps_teacher_forced(G, tag, out, args, beta)
    assert n_exempt <= 0.02 * n_rows
    # whole run: a relu'/arg-min flip in one guided step moves that row by up to lr per later guided step, so the bulk
    # is held to 1e-5 and every row to lr * (guided steps)
    a, b = npy(out["final_iterate"]), G[tag + "|final_iterate"]
    err = (np.abs(a - b) / np.array([0.5, 5.0])).reshape(a.shape[0], -1).max(axis=1)
    n_guided = int(G[tag + "|n_guidance_calls"])
    assert np.median(err) < 1e-5 and err.max() < args.guidance_lr * n_guided, (np.median(err), err.max())


def test_refinement_golden(golden_dir):
    """--refinement (reference nusc_train.py:1034-1071): 50 Adam steps on the mixing logits of the violating rows"""
    G = np.load(os.path.join(golden_dir, "sampler_modes.npz"))
    out, net, batch, args = _pipeline(NT.OURS_FLAGS + ["--refinement"], 2013, 2)
    assert out.get("refinement")
    a, b, base = npy(out["controls"]), G["refine|final_controls"], G["refine|controls"]
    moved = np.abs(b - base).reshape(b.shape[0], -1).max(axis=1) > 1e-6
    assert moved.any()
    # rows the reference left alone are left alone: the same RefineNet output
    close(a[~moved], b[~moved], what="untouched rows")
    # Teacher-forced parity of the optimisation (the reference's own logits going into iterations 0, 1, 2, 10, 25, 49):
    # the scores of that iteration's mixed controls and the gradient Adam is fed
    mix = NT.MixingProblem(out["pack"], torch.from_numpy(base).cuda(), out["iterates"], NT.build_stl_cache(args), args)
    close_elem(mix.scores0, G["refine|mix_scores0"], rtol=2e-5, floor=1.0, what="scores before mixing")
    for j, it in enumerate(G["refine|mix_iters"]):
        lam = torch.from_numpy(G["refine|mix_lam_in"][j]).cuda().requires_grad_()
        with torch.enable_grad():
            _, sc = mix.backward_into(lam)
        sref, gref = G["refine|mix_scores"][j], G["refine|mix_grad"][j]
        close_elem(sc, sref, rtol=5e-5, floor=1.0, what="mix scores, iteration %d" % it)
        edge = np.abs(5e-4 - sref) < 1e-4  # relu'(5e-4 - score) flips within the score tolerance: rows exempt, counted
        assert edge.mean() < 0.05
        scale = np.abs(gref).max(axis=1, keepdims=True)
        ok = ~edge
        rel = (np.abs(npy(lam.grad) - gref) / np.maximum(scale, 1e-12))[ok]
        # early iterations: 1e-3 of the row's largest entry.  Later the logits are saturated and the softmax backward
        # r (g - sum r g) cancels to a small remainder of large terms (summed in another order on the device): bulk 1e-3
        assert rel.max() <= 1e-3 if it <= 2 else (np.percentile(rel, 99) <= 1e-3 and rel.max() <= 0.1), ("grad", it, rel.max())
    # the first Adam step from the shared start (all-ones logits): lr * g / (|g| + 1e-8)
    lam = torch.from_numpy(G["refine|mix_lam_in"][0]).cuda().requires_grad_()
    opt = torch.optim.Adam([lam], lr=3e-1)
    with torch.enable_grad():
        mix.backward_into(lam)
    opt.step()
    g0 = np.abs(G["refine|mix_grad"][0])
    firm = (g0 == 0) | (g0 > 1e-6)  # entries whose step does not hinge on |g| ~ Adam's eps
    assert firm.mean() > 0.9
    assert np.abs(npy(lam) - G["refine|mix_lam_out"][0])[firm].max() < 1e-4
    # The 50-iteration trajectory itself is chaotic at lr 0.3 (logits saturate, rows pick one iterate): the end state is
    # held to the success rate, not to the controls
    sa, sb = npy(out["scores"]), G["refine|scores"]
    assert abs((sa > 0).mean() - (sb > 0).mean()) <= 0.03
    err = (np.abs(a - b) / np.array([0.5, 5.0])).reshape(a.shape[0], -1).max(axis=1)
    assert np.median(err[moved]) < 1e-2, np.median(err[moved])


def test_mono_rollout_golden(golden_dir):
    """diffusion_rollout(mono=True) in the --gt_data_training layout (reference nusc_train.py:570-572)"""
    G = np.load(os.path.join(golden_dir, "sampler_modes.npz"))
    S_, bs, nt, seed = 16, 2, 20, 2014
    args = NT.default_args(n_randoms=S_, sampling_size=S_)
    batch = cuda(synthetic.make_scene_batch(bs, nt=nt, n_randoms=S_, seed=seed))
    net = Net(args)
    net.load_state_dict(synthetic.make_weights(1007, nt=nt), strict=True)
    net = net.cuda()
    n = bs * S_
    args.inject_noise = [t.cuda() for t in synthetic.noise_stream(seed + 77, n, nt * 2, 99)]
    args.keep_all_iterates = True
    with torch.no_grad():
        feature = net.encode_feat(batch)
    gt_stlp = batch["pre_stlp"].reshape(bs, S_, 3, 6)[:, 0, 0]
    res = NT.diffusion_rollout(torch.zeros(n, nt * 2, device="cuda"), net, batch, batch["gt_high_level"], feature, args,
                               NT.get_diffusion_coeffs(args), mono=True, tmp_stlp=gt_stlp)
    close(res[0], G["mono|final"], what="final")
    close(res[1][50], G["mono|iter_mid"], what="iterate 50")
    # entry 0 of final_list is x_T itself (normalised), as upstream's res_list[0]
    close(res[1][0], NT.normalize_diff(args.inject_noise[0], n, nt, 0.5, 5.0, True))


def test_fastforward_returns_initial_noise():
    """fastforward=True skips the reverse loop (reference nusc_train.py:567): the result is normalize_diff(x_T)"""
    S_, bs, nt = 16, 2, 20
    args = NT.default_args(n_randoms=S_, sampling_size=S_)
    n = bs * S_ * 3
    x_T = torch.randn(n, nt * 2, device="cuda")
    args.inject_noise = [x_T]
    net = Net(args).cuda()
    b = cuda(synthetic.make_scene_batch(bs, nt=nt, n_randoms=S_, seed=1))
    final, fl = NT.diffusion_rollout(torch.zeros(n, nt * 2, device="cuda"), net, b, None, None, args,
                                     NT.get_diffusion_coeffs(args), fastforward=True)
    want = NT.normalize_diff(x_T, n, nt, args.mul_w_max, args.mul_a_max, args.diffusion_clip)
    assert torch.equal(final, want) and len(fl) == 1 and torch.equal(fl[-1], want)


# ------------------------------------------------------------------------------------------
# RefineNet variants
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,extra,ex_in", [("cat", ["--diverse_loss", "--diverse_fuse_type", "cat"], 40),
                                             ("noarch", ["--diverse_loss", "--no_arch"], 0), ("plain", [], 0),
                                             ("clip", ["--diverse_loss", "--clip_rect"], 0)])
def test_rect_variants_golden(golden_dir, tag, extra, ex_in):
    G = np.load(os.path.join(golden_dir, "rect_variants.npz"))
    bs, S_, nt, seed = 2, 16, 20, 2015
    base = ["-e", "x", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--rect_head", "--flex", "--skip_nusc_load"]
    args = NT.default_args(base + extra, n_randoms=S_, sampling_size=S_)
    net = Net(args)
    sd = synthetic.make_weights(1007, nt=nt, rect_extra_in=ex_in)
    if not args.diverse_loss:
        sd = {k: v for k, v in sd.items() if not k.startswith("merge_net")}
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    n = bs * S_ * 3
    batch = cuda(synthetic.make_scene_batch(bs, nt=nt, n_randoms=S_, seed=seed))
    with torch.no_grad():
        feat = net.encode_feat(batch).reshape(bs, 1, -1).repeat(1, S_ * 3, 1).reshape(n, -1)
        hl = torch.tensor([0.0, 1.0, 2.0], device="cuda").repeat(bs * S_).reshape(n, 1)
        out = net.rect_forward(feat, hl, batch["pre_stlp"].reshape(n, 6), torch.from_numpy(G[tag + "|u0"]).cuda(),
                               torch.from_numpy(G[tag + "|scores"]).cuda())
    close_elem(out, G[tag + "|out"], rtol=2e-5, floor=1.0, what=tag)


# ------------------------------------------------------------------------------------------
# the benchmarked path: bf16 tcgen05 denoiser + RefineNet, against the reference's golden run
# ------------------------------------------------------------------------------------------
BF16_CTRL_TOL = 2e-2     # north_star: normalised control units (measured on this fixture: 1.1e-3 iterates, 2.9e-3 refined)
BF16_SCORE_TOL = 0.1     # robustness units: the bf16-induced score bound (measured max 0.040 on the candidate scores,
                         # 0.071 on the final ones); top-2 margins below 2x it are not decided by bf16 arithmetic


def test_pipeline_bf16_golden(golden_dir):
    """README "Ours" on the tcgen05 engine (the path bench.py times) against pipeline.npz of the unmodified reference:
    iterates and refined controls within 2e-2 normalised, candidate / final scores within BF16_SCORE_TOL, and the
    selected candidate equal to the reference's wherever its top-2 margin exceeds 2 * BF16_SCORE_TOL."""
    G = np.load(os.path.join(golden_dir, "pipeline.npz"))
    out, net, batch, args = _pipeline(NT.OURS_FLAGS, 2001, 2, precision="bf16")
    scale = np.array([0.5, 5.0])
    e_it = np.abs(npy(out["final_iterate"]) - G["ours|final_iterate"]) / scale
    assert e_it.max() < BF16_CTRL_TOL, e_it.max()
    cs, cref = npy(out["cand_scores"]), G["ours|cand_scores"]
    assert np.abs(cs - cref).max() < BF16_SCORE_TOL, np.abs(cs - cref).max()
    srt = np.sort(cref, axis=0)
    clear = (srt[-1] - srt[-2]) > 2 * BF16_SCORE_TOL
    assert clear.sum() > 0.3 * clear.size  # the check is not vacuous
    assert (npy(out["best_idx"])[clear] == cref.argmax(0)[clear]).all()
    same = npy(out["best_idx"]) == cref.argmax(0)
    e_c = (np.abs(npy(out["controls"]) - G["ours|controls"]) / scale).reshape(len(same), -1).max(axis=1)
    assert e_c[same].max() < 2 * BF16_CTRL_TOL, e_c[same].max()  # RefineNet head adds its own bf16 pass
    e_s = np.abs(npy(out["scores"]) - G["ours|scores"])
    assert e_s[same].max() < 1.5 * BF16_SCORE_TOL, e_s[same].max()
    print("bf16 vs reference: iterate %.2e  cand scores %.2e  controls %.2e  scores %.2e  decided rows %d/%d"
          % (e_it.max(), np.abs(cs - cref).max(), e_c[same].max(), e_s[same].max(), clear.sum(), clear.size))


# ------------------------------------------------------------------------------------------
# handle refresh after in-place parameter updates; captured guidance
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_handle_refresh_tracks_inplace_updates(precision):
    S_, bs, nt = 64, 2, 20
    args = NT.default_args(n_randoms=S_, sampling_size=S_, precision=precision)
    net = Net(args)
    net.load_state_dict(synthetic.make_weights(1007, nt=nt))
    net = net.cuda()
    b = cuda(synthetic.make_scene_batch(bs, nt=nt, n_randoms=S_, seed=9))
    args.inject_noise = [t.cuda() for t in synthetic.noise_stream(5, bs * S_ * 3, nt * 2, 99)]
    stls, co = NT.build_stl_cache(args), NT.get_diffusion_coeffs(args)
    o1 = NT.sample_and_score(net, b, stls, co, args)["controls"].clone()
    h1 = net.native_handle(precision).value
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.01)  # in place: same storage, new version
    o2 = NT.sample_and_score(net, b, stls, co, args)["controls"].clone()
    assert net.native_handle(precision).value == h1  # refreshed, not rebuilt
    fresh = Net(args)
    fresh.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()})
    fresh = fresh.cuda()
    o3 = NT.sample_and_score(fresh, b, stls, co, args)["controls"]
    assert torch.equal(o2, o3) and not torch.equal(o1, o2)
    # edits through .data do not bump the version: invalidate_native() is the documented hook
    net.policy_net[4].bias.data.add_(0.05)
    net.invalidate_native()
    o4 = NT.sample_and_score(net, b, stls, co, args)["controls"]
    assert not torch.equal(o4, o2)


def test_captured_pipeline_with_guidance():
    """--guidance under one CUDA graph: the batch normaliser is a device word, so the replay follows the batch"""
    S_ = 64
    args = NT.default_args(NT.GUIDANCE_FLAGS, precision="bf16", n_randoms=S_, sampling_size=S_)
    net = Net(args)
    net.load_state_dict(synthetic.make_weights(1007))
    net = net.cuda()
    stls, co = NT.build_stl_cache(args), NT.get_diffusion_coeffs(args)
    progs = NT._fused_programs(stls, args.nt)
    b1 = cuda(synthetic.make_scene_batch(4, n_randoms=S_, seed=41))
    b2 = cuda(synthetic.make_scene_batch(4, n_randoms=S_, seed=42, lane_valid_p=0.2))  # another validity mean
    runner = NT.CapturedPipeline(net, stls, co, args, b1)
    for b in (b1, b2):
        o = {k: v.clone() for k, v in runner(b).items() if isinstance(v, torch.Tensor)}
        nb = NT.LazyBatch({k: b[k] for k in ("ego_traj", "neighbors", "currlane_wpts", "leftlane_wpts", "rightlane_wpts",
                                             "curr_id", "left_id", "right_id", "gt_high_level", "pre_stlp")})
        nb["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
        pack = NT.augment_batch_data(nb, None, args, n_randoms=S_)["_pstl_pack"]
        assert torch.equal(NT.score_pack(pack, o["controls"], args, progs)["best_score"], o["scores"])
        assert torch.isfinite(o["scores"]).all()
    # guidance raises the share of satisfied rows over the unguided run on the same batch (fresh noise, so a loose test)
    a0 = NT.default_args(NT.GUIDANCE_FLAGS, precision="bf16", n_randoms=S_, sampling_size=S_, guidance=False)
    acc0 = float(NT.sample_and_score(net, b1, stls, co, a0)["acc"])
    acc1 = float(runner(b1)["acc"])
    assert acc1 >= acc0 - 0.02, (acc0, acc1)


def test_fused_programs_follow_the_formulas_not_the_first_object():
    """the compiled programs used to be cached on stls_cac[0] keyed by (T, device) only, so a second list that
    shared its first formula — or a ListAnd edited in place — silently reused the first list's programs."""
    args = NT.default_args()
    a = NT.build_stl_cache(args)
    pa = NT._fused_programs(a, 20)
    b = [a[0], a[2], a[1]]  # same first formula, the other two swapped
    pb = NT._fused_programs(b, 20)
    assert pb[0] is pa[0] and pb[1] is pa[2] and pb[2] is pa[1]
    assert NT._fused_programs(a, 20)[1] is pa[1]
    if hasattr(a[1], "lists") and len(a[1].lists) > 1:
        import copy
        c = [a[0], copy.copy(a[1]), a[2]]
        c[1].lists = list(a[1].lists[:-1])  # an edited ListAnd compiles to another program
        assert NT._fused_programs(c, 20)[1] is not pa[1]
