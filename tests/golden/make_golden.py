"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only:   python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle_golden.py) and are the golden outputs the CUDA
path is compared with on the GPU box, where /root/reference does not exist.  Inputs are not
stored: they are regenerated from seeds by pstl_b200.synthetic (same torch build on both sides);
every fixture carries input checksums so a drifting generator is detected, not silently used.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from formulas import recipes  # noqa: E402
import pstl_b200  # noqa: E402
from pstl_b200 import synthetic  # noqa: E402


def checksum(t):
    return float(t.double().abs().sum())


def kat_inputs(seed=11, N=6, T=12):
    g = torch.Generator().manual_seed(seed)
    x = {k: (torch.rand(N, T, generator=g) * 2 - 1) for k in "abc"}
    # adversarial rows: exact ties, large magnitudes, identical signals
    x["a"][0, 3] = x["a"][0, 4]
    x["b"][1] = x["a"][1]
    x["c"][2] = x["c"][2] * 30
    x["a"][3] = 0.0
    return x


def gen_stl_kats():
    import stl_d_lib as R
    out = {}
    x0 = kat_inputs()
    out["in_checksum"] = np.array([checksum(x0[k]) for k in "abc"])
    fs = recipes(R)
    names = sorted(fs)
    out["names"] = np.array(names)
    for name in names:
        for tau in (1.0, 100.0):
            for hard in (False, True):
                x = {k: v.clone().requires_grad_() for k, v in x0.items()}
                d = {"hard": True} if hard else None
                y = fs[name](x, tau, d)
                key = "%s|%g|%d" % (name, tau, int(hard))
                out[key] = y.detach().numpy()
                if not hard and torch.isfinite(y[:, 0]).all():
                    (gr,) = torch.autograd.grad(y[:, 0].sum(), [x["a"]], allow_unused=True, retain_graph=True)
                    grs = torch.autograd.grad(y[:, 0].sum(), [x["a"], x["b"], x["c"]], allow_unused=True)
                    for k, gk in zip("abc", grs):
                        out[key + "|g" + k] = (torch.zeros_like(x0[k]) if gk is None else gk).numpy()
    # the survey's seed KATs, T=8
    a = torch.tensor([[.30, -.20, .50, .10, -.40, .25, .05, .60]])
    b = torch.tensor([[-.10, .40, .20, -.30, .35, .15, -.05, .45]])
    xs = {"a": a, "b": b, "c": a}
    for name in ("always_0_3", "eventually_1_4", "and", "or", "imply", "until_0_8", "until_2_5", "once_m3_0",
                 "ev_alw_and"):
        out["seed|" + name] = fs[name](xs, 100.0).numpy()
    out["seed|ev_alw_and|tau1"] = fs["ev_alw_and"](xs, 1.0).numpy()
    out["str_symbol"] = np.array(str(fs["ev_alw_and"]))
    fs["ev_alw_and"].update_format("word")
    out["str_word"] = np.array(str(fs["ev_alw_and"]))
    np.savez_compressed(os.path.join(HERE, "stl_kats.npz"), **out)
    print("stl_kats:", len(out), "arrays")


def gen_dense(T, args):
    """compute_stl_dense on config-1-shaped dense rows, plus the predicate signals and the
    gradient of the guidance-style loss w.r.t. the ego trajectory."""
    out = {}
    for tag, n, nt, knei, seed in (("t20k8", 192, 20, 8, 1008), ("t50k16", 48, 50, 16, 1009)):
        args.nt = nt
        stls = T.build_stl_cache(args)
        x, idx, mask = synthetic.make_dense_stl_input(n, nt=nt, n_neighbors=knei, seed=seed)
        out[tag + "|in_checksum"] = np.array([checksum(x[k]) for k in sorted(x)])
        x["ego_traj"] = x["ego_traj"].clone().requires_grad_()
        scores_list, scores, acc, xo = T.compute_stl_dense(x, stls, idx, mask, args, debug=True)
        for k in ("x2curr_d", "x2curr_th", "x2left_d", "x2left_th", "x2right_d", "x2right_th", "min_nei_d"):
            out[tag + "|" + k] = xo[k].detach().numpy()
        out[tag + "|scores"] = scores.detach().numpy()
        out[tag + "|scores3"] = torch.stack(scores_list[:3], 0).detach().numpy()
        out[tag + "|acc"] = np.array(acc.item())
        loss = T.mask_mean(torch.relu(args.stl_nn_thres - scores), mask)
        (g,) = torch.autograd.grad(loss, [x["ego_traj"]])
        out[tag + "|grad_ego"] = g.numpy()
    args.nt = 20
    np.savez_compressed(os.path.join(HERE, "stl_dense.npz"), **out)
    print("stl_dense:", len(out), "arrays")


def run_pipeline(flags, tag, bs, S, seed, out):
    """Drive the reference's own run_sampling_test with a one-batch fake loader, injecting the
    seeded noise stream and capturing the intermediates of the timed region."""
    import importlib
    T, args = ref_shim.load(flags + ["--n_randoms", str(S), "--sampling_size", str(S), "--n_trials", "0"])
    import nusc_model
    nt = args.nt
    batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
    sd = synthetic.make_weights(seed=1007, nt=nt)
    net = nusc_model.Net(args)
    missing = net.load_state_dict(sd, strict=True)
    stls = T.build_stl_cache(args)
    coeffs = T.get_diffusion_coeffs(args)
    N = bs * S * 3
    stream = synthetic.noise_stream(seed + 77, N, nt * 2, args.diffusion_steps - 1)
    it = iter(stream)
    cap = {"stl": [], "rect": []}
    real_randn_like = torch.randn_like
    torch.randn_like = lambda t, **k: next(it).to(t.dtype)
    real_roll, real_stl, real_rect = T.diffusion_rollout, T.compute_stl_dense, net.rect_forward
    real_gen = T.generate_trajs
    cap["us"] = []

    def gen(s_, us_, dt_):
        cap["us"].append(us_.detach().clone())
        if len(cap["us"]) > 2:
            del cap["us"][0]
        return real_gen(s_, us_, dt_)

    T.generate_trajs = gen
    # guided reverse steps: what every fresh Adam of nusc_train.py:607 saw and did (mu before, gradient, mu after)
    real_adam = torch.optim.Adam
    cap["gsteps"] = []

    class RecordingAdam(real_adam):
        def step(self, *a, **k):
            p = self.param_groups[0]["params"][0]
            rec = [p.detach().clone(), p.grad.detach().clone()]
            r = super().step(*a, **k)
            rec.append(p.detach().clone())
            cap["gsteps"].append(rec)
            return r

    if args.guidance or getattr(args, "refinement", False):
        torch.optim.Adam = RecordingAdam

    def roll(*a, **k):
        r = real_roll(*a, **k)
        cap["rollout"] = r
        return r

    def stl(*a, **k):
        r = real_stl(*a, **k)
        cap["stl"].append(r[1].detach().clone())
        return r

    def rect(*a, **k):
        r = real_rect(*a, **k)
        cap["rect"].append(r.detach().clone())
        return r

    T.diffusion_rollout, T.compute_stl_dense, net.rect_forward = roll, stl, rect
    real_md = T.napi.measure_diversity
    real_ex = T.napi.measure_extra_diversity
    z = torch.zeros(())
    T.napi.measure_diversity = lambda *a, **k: (z, z, [np.zeros(1)], [np.zeros(1)])
    T.napi.measure_extra_diversity = lambda *a, **k: {}
    try:
        import io
        import contextlib
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            try:
                T.run_sampling_test(stls, [batch], net, coeffs, args, None, None)
            except KeyError:
                # the print line reads meters our stubbed diversity metrics did not fill;
                # everything inside the timed region has already run and been captured
                pass
    finally:
        torch.randn_like = real_randn_like
        T.diffusion_rollout, T.compute_stl_dense = real_roll, real_stl
        T.generate_trajs = real_gen
        torch.optim.Adam = real_adam
        T.napi.measure_diversity, T.napi.measure_extra_diversity = real_md, real_ex
    final, feature, its = cap["rollout"]
    K = args.multi_cands
    out[tag + "|in_checksum"] = np.array([checksum(batch[k]) for k in sorted(batch)] + [checksum(stream[0]),
                                         sum(checksum(v) for v in sd.values())])
    out[tag + "|feature"] = feature.reshape(bs, S * 3, -1)[:, 0].detach().numpy()
    out[tag + "|final_iterate"] = final.detach().numpy()
    out[tag + "|iter_mid"] = its[50].detach().numpy()
    # stl calls in order: [0]=traj-opt reference rows, [guidance calls...], best-of-K, n_rolls..., final
    n_ref = 51 if getattr(args, "refinement", False) else 0  # --refinement: one scoring + 50 optimisation iterations
    n_guid = len(cap["stl"]) - 1 - 1 - (args.n_rolls or 0) - 1 - n_ref
    out[tag + "|n_guidance_calls"] = np.array(n_guid)
    out[tag + "|tj_scores"] = cap["stl"][0].numpy()
    out[tag + "|cand_scores"] = cap["stl"][1 + n_guid].reshape(K, N).numpy()
    out[tag + "|rect_first"] = cap["rect"][0].numpy()
    out[tag + "|controls"] = cap["rect"][-1].numpy()
    out[tag + "|final_controls"] = cap["us"][-1].numpy()  # what the last generate_trajs rolled out (differs under --refinement)
    out[tag + "|scores"] = cap["stl"][-1].numpy()
    if args.guidance and args.guidance_niters == 1:
        assert len(cap["gsteps"]) == n_guid
        out[tag + "|gstep_mu_in"] = torch.stack([r[0] for r in cap["gsteps"]]).reshape(n_guid, N, -1).numpy()
        out[tag + "|gstep_grad"] = torch.stack([r[1] for r in cap["gsteps"]]).reshape(n_guid, N, -1).numpy()
        out[tag + "|gstep_mu_out"] = torch.stack([r[2] for r in cap["gsteps"]]).reshape(n_guid, N, -1).numpy()
        out[tag + "|gstep_scores"] = torch.stack(cap["stl"][1:1 + n_guid]).numpy()
    if getattr(args, "refinement", False):
        # the mixing logits of --refinement (nusc_train.py:1034-1071) at a few of the 50 Adam iterations: value going in,
        # gradient, value coming out, and the scores of that iteration's mixed controls
        assert len(cap["gsteps"]) == 50
        its_kept = [0, 1, 2, 10, 25, 49]
        out[tag + "|mix_iters"] = np.array(its_kept)
        out[tag + "|mix_lam_in"] = torch.stack([cap["gsteps"][j][0] for j in its_kept]).numpy()
        out[tag + "|mix_grad"] = torch.stack([cap["gsteps"][j][1] for j in its_kept]).numpy()
        out[tag + "|mix_lam_out"] = torch.stack([cap["gsteps"][j][2] for j in its_kept]).numpy()
        out[tag + "|mix_scores"] = torch.stack([cap["stl"][-51 + j] for j in its_kept]).numpy()
        out[tag + "|mix_scores0"] = cap["stl"][-52].numpy()
    print(tag, "N=%d stl_calls=%d rect_calls=%d guidance_calls=%d" % (N, len(cap["stl"]), len(cap["rect"]), n_guid))


def gen_pipeline():
    out = {}
    run_pipeline(ref_shim.OURS_FLAGS, "ours", bs=2, S=16, seed=2001, out=out)
    run_pipeline(ref_shim.GUIDE_FLAGS, "guide", bs=2, S=16, seed=2002, out=out)
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **out)


TRAJOPT_FLAGS = ["-e", "e1_trajopt", "--diffusion", "--load_stlp", "--flex", "--skip_nusc_load", "--trajopt_only"]


def gen_trajopt(iters=15, bs=2, S=16, seed=2003):
    """The reference's trajectory-optimisation loop (nusc_train.py:1303-1325): generate_trajs ->
    pre_prepare_stl_cache -> compute_trajopt_loss_lite -> Adam(lr=trajopt_lr), driven on one synthetic batch."""
    T, args = ref_shim.load(TRAJOPT_FLAGS + ["--n_randoms", str(S)])
    nt = args.nt
    batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
    out = {"in_checksum": np.array([checksum(batch[k]) for k in sorted(batch)])}
    stls = T.build_stl_cache(args)
    b = {k: v.clone() for k, v in batch.items()}
    b["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
    gt_stlp = b["pre_stlp"].reshape(bs, S, 3, 6)[:, 0, 0]  # any (bs,6): --load_stlp takes the dense pSTL from pre_stlp
    b = T.augment_batch_data(b, gt_stlp, args)
    states = b["ego_traj"][:, 0, :4]
    dense_states = states.unsqueeze(1).unsqueeze(1).repeat(1, S, 3, 1)
    real_md = T.napi.measure_diversity
    z = torch.zeros(())
    T.napi.measure_diversity = lambda *a, **k: (z, z, [np.zeros(1)], [np.zeros(1)])
    try:
        params = b["params"] = b["params"].clone().requires_grad_()
        opt = torch.optim.Adam([params], lr=args.trajopt_lr)
        for ii in range(iters):
            trajs = T.generate_trajs(dense_states, params, args.dt)
            cache = T.pre_prepare_stl_cache(b)
            res = T.compute_trajopt_loss_lite(params, trajs, stls, cache, ii, iters)
            loss, dense_loss, reg_loss, dense_scores = res[0], res[1], res[2], res[5]
            out["loss|%d" % ii] = np.array([loss.item(), dense_loss.item(), reg_loss.item()])
            if ii in (0, iters - 1):
                out["scores|%d" % ii] = dense_scores.detach().reshape(-1).numpy().copy()
            opt.zero_grad()
            loss.backward()
            if ii == 0:
                out["grad|0"] = params.grad.detach().reshape(-1, nt, 2).numpy().copy()
            if ii in (4, 9):  # one mid-run iteration in full, for the teacher-forced check of a single step
                st = opt.state[params]
                out["tf%d|params_in" % ii] = params.detach().reshape(-1, nt, 2).numpy().copy()
                out["tf%d|m" % ii] = st["exp_avg"].reshape(-1, nt, 2).numpy().copy()
                out["tf%d|v" % ii] = st["exp_avg_sq"].reshape(-1, nt, 2).numpy().copy()
                out["tf%d|grad" % ii] = params.grad.detach().reshape(-1, nt, 2).numpy().copy()
                out["tf%d|scores" % ii] = dense_scores.detach().reshape(-1).numpy().copy()
            opt.step()
            if ii in (4, 9):
                out["tf%d|params_out" % ii] = params.detach().reshape(-1, nt, 2).numpy().copy()
            if ii in (0, 4, iters - 1):
                out["params|%d" % ii] = params.detach().reshape(-1, nt, 2).numpy().copy()
    finally:
        T.napi.measure_diversity = real_md
    out["hyper"] = np.array([args.trajopt_lr, args.stl_trajopt_thres, args.reg_loss, args.mul_w_max, args.mul_a_max, iters])
    np.savez_compressed(os.path.join(HERE, "trajopt.npz"), **out)
    print("trajopt:", len(out), "arrays; loss", out["loss|0"], "->", out["loss|%d" % (iters - 1)])


TRAIN_FLAGS = ["-e", "e7_ours", "--diffusion", "--load_stlp", "--rect_head", "--flex", "--multi_cands", "5",
               "--skip_nusc_load"]  # README "Ours" training command minus the loss switches varied below
LOSS_VARIANTS = {
    "ours": ["--stl_weight", "0.0", "--diverse_loss"],
    "weighted": ["--stl_weight", "0.7", "--diverse_loss", "--rect_reg_loss", "0.3", "--diversity_scale", "0.8",
                 "--diversity_weight", "1.5", "--stl_nn_thres", "0.05"],
    "detach": ["--stl_weight", "0.5", "--diverse_loss", "--diverse_detach", "--rect_reg_loss", "0.2", "--n_shards", "2"],
    "plain": ["--stl_weight", "0.5", "--rect_reg_loss", "0.4", "--extra_rect_reg", "0.6"],
}


def gen_losses(bs=3, S=16, seed=2005, warm_iters=120):
    """compute_policy_loss (nusc_train.py:370-478) of the --rect_head training step on one synthetic batch, for the
    README flags and three variants: the loss terms, the scores it derived, d loss / d rect_controls (total, through
    the STL scores), the same with the trajectories detached (the direct part) and d loss / d scores.  The controls
    come out of a short run of the reference's own traj-opt so that a fair share of the rows satisfies its formula."""
    out = {}
    T, args = ref_shim.load(TRAJOPT_FLAGS + ["--n_randoms", str(S), "--trajopt_lr", "0.03"])
    nt = args.nt
    batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
    out["in_checksum"] = np.array([checksum(batch[k]) for k in sorted(batch)])
    N = bs * S * 3

    def prepare(T, args):
        b = {k: v.clone() for k, v in batch.items()}
        b["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
        gt_stlp = b["pre_stlp"].reshape(bs, S, 3, 6)[:, 0, 0]
        b = T.augment_batch_data(b, gt_stlp, args)
        st = b["ego_traj"][:, 0, :4]
        return b, st.unsqueeze(1).unsqueeze(1).repeat(1, S, 3, 1)

    real_md = T.napi.measure_diversity
    z = torch.zeros(())
    T.napi.measure_diversity = lambda *a, **k: (z, z, [np.zeros(1)], [np.zeros(1)])
    try:
        b, dense_states = prepare(T, args)
        stls = T.build_stl_cache(args)
        params = b["params"].clone().requires_grad_()
        opt = torch.optim.Adam([params], lr=args.trajopt_lr)
        for ii in range(warm_iters):
            trajs = T.generate_trajs(dense_states, params, args.dt)
            res = T.compute_trajopt_loss_lite(params, trajs, stls, T.pre_prepare_stl_cache(b), ii, warm_iters)
            opt.zero_grad()
            res[0].backward()
            opt.step()
        nn_controls = params.detach().reshape(N, nt, 2).clone()
        g = torch.Generator().manual_seed(seed)
        lim = torch.tensor([args.mul_w_max, args.mul_a_max])
        rect0 = nn_controls + 0.05 * lim * torch.randn(N, nt, 2, generator=g)
        rect0[::7] *= 1.6  # some rows leave the control box
        out["nn_controls"] = nn_controls.numpy()
        out["rect_controls"] = rect0.numpy()
        for tag, extra in LOSS_VARIANTS.items():
            T, args = ref_shim.load(TRAIN_FLAGS + extra + ["--n_randoms", str(S)])
            b, dense_states = prepare(T, args)
            flat_states = dense_states.reshape(N, 4)
            stls = T.build_stl_cache(args)
            nn_trajs = T.generate_trajs(flat_states, nn_controls, args.dt)
            zeros = torch.zeros(N, nt * 2)

            def run(detach_trajs):
                rect = rect0.clone().requires_grad_()
                rect_trajs = T.generate_trajs(flat_states, rect, args.dt)
                if detach_trajs:
                    rect_trajs = rect_trajs.detach()
                extras = (None, zeros, b["highlevel_dense"], torch.zeros(N), b["valids_dense"].reshape(-1), 0, zeros,
                          nn_controls, None, rect)
                rd, _ = T.compute_policy_loss(b, None, stls, nn_trajs, rect_trajs, None, args, diffusion_extras=extras)
                return rect, rd

            rect, rd = run(False)
            gs = [rect] + ([rd["scores"]] if rd["scores"].requires_grad else [])
            grads = torch.autograd.grad(rd["loss"], gs, allow_unused=True)
            out[tag + "|scores"] = rd["scores"].detach().numpy()
            out[tag + "|losses"] = np.array([float(rd[k].detach()) if k in rd else np.nan for k in
                                             ("loss", "loss_stl", "loss_reg", "loss_diversity", "extra_loss_reg")])
            out[tag + "|grad_total"] = grads[0].numpy()
            out[tag + "|grad_scores"] = (grads[1] if len(grads) > 1 and grads[1] is not None else torch.zeros(N)).numpy()
            rect, rd = run(True)
            (gd,) = torch.autograd.grad(rd["loss"], [rect], allow_unused=True)
            out[tag + "|grad_direct"] = (gd if gd is not None else torch.zeros_like(rect)).numpy()
            out[tag + "|hyper"] = np.array([args.stl_nn_thres, args.stl_weight, args.diversity_scale,
                                            args.diversity_weight, args.rect_reg_loss, args.extra_rect_reg or 0.0,
                                            args.n_shards, int(args.diverse_loss), int(args.diverse_detach),
                                            args.mul_w_max, args.mul_a_max])
            print("losses[%s]:" % tag, out[tag + "|losses"], "accepted %.2f" % float((rd["scores"] > 0).float().mean()))
    finally:
        T.napi.measure_diversity = real_md
    out["valid"] = b["valids_dense"].reshape(-1).numpy()
    out["shape"] = np.array([bs, S, nt])
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)
    print("losses:", len(out), "arrays")


def gen_refine_step(bs=3, S=16, seed=2005):
    """One --rect_head training step of the reference (nusc_train.py:1402-1405 rect_forward, :1420-1427 the loss,
    :1523-1525 Adam over net.rect_net.parameters()) on the controls of losses.npz: RefineNet output, loss terms,
    d loss / d rect_controls, the six rect_net gradients and the parameters after the step."""
    Lz = np.load(os.path.join(HERE, "losses.npz"))
    T, args = ref_shim.load(TRAIN_FLAGS + LOSS_VARIANTS["weighted"] + ["--n_randoms", str(S)])
    import nusc_model
    nt = args.nt
    N = bs * S * 3
    batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
    b = {k: v.clone() for k, v in batch.items()}
    b["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
    b = T.augment_batch_data(b, b["pre_stlp"].reshape(bs, S, 3, 6)[:, 0, 0], args)
    flat_states = b["ego_traj"][:, 0, :4].unsqueeze(1).repeat(1, S * 3, 1).reshape(N, 4)
    stls = T.build_stl_cache(args)
    net = nusc_model.Net(args)
    net.load_state_dict(synthetic.make_weights(seed=1007, nt=nt), strict=True)
    g = torch.Generator().manual_seed(seed + 1)
    feat_scene = 0.5 * torch.randn(bs, 224, generator=g)
    feature = feat_scene.unsqueeze(1).repeat(1, S * 3, 1).reshape(N, 224)
    nn_controls = torch.from_numpy(Lz["nn_controls"])
    out = {"feat_scene": feat_scene.numpy()}
    real_md = T.napi.measure_diversity
    z = torch.zeros(())
    T.napi.measure_diversity = lambda *a, **k: (z, z, [np.zeros(1)], [np.zeros(1)])
    try:
        prev_trajs = T.generate_trajs(flat_states, nn_controls, args.dt)
        prev_in = T.pre_prepare_stl_cache(b, dense_trajs=prev_trajs[:, :-1])
        _, prev_scores, _ = T.compute_stl_dense(prev_in, stls, b["highlevel_dense"], prev_in["dense_valids"].reshape(-1), args)
        opt = torch.optim.Adam(net.rect_net.parameters(), lr=args.lr)
        rect = net.rect_forward(feature, b["highlevel_dense"], b["stlp_dense"][:, 0], nn_controls.detach(), prev_scores.detach())
        rect.retain_grad()
        rect_trajs = T.generate_trajs(flat_states, rect, args.dt)
        zeros = torch.zeros(N, nt * 2)
        extras = (None, zeros, b["highlevel_dense"], torch.zeros(N), b["valids_dense"].reshape(-1), 0, zeros, nn_controls, None, rect)
        rd, _ = T.compute_policy_loss(b, None, stls, prev_trajs, rect_trajs, None, args, diffusion_extras=extras)
        opt.zero_grad()
        rd["loss"].backward()
        out["prev_scores"] = prev_scores.detach().numpy()
        out["rect"] = rect.detach().numpy()
        out["grad_rect"] = rect.grad.numpy()
        out["losses"] = np.array([float(rd[k].detach()) for k in ("loss", "loss_stl", "loss_reg", "loss_diversity")])
        for li in (0, 2, 4):
            out["g_w%d" % li] = net.rect_net[li].weight.grad.numpy().copy()
            out["g_b%d" % li] = net.rect_net[li].bias.grad.numpy().copy()
        opt.step()
        for li in (0, 2, 4):
            out["w%d_after" % li] = net.rect_net[li].weight.detach().numpy().copy()
            out["b%d_after" % li] = net.rect_net[li].bias.detach().numpy().copy()
        out["lr"] = np.array(args.lr)
    finally:
        T.napi.measure_diversity = real_md
    np.savez_compressed(os.path.join(HERE, "refine_step.npz"), **out)
    print("refine_step:", len(out), "arrays; loss", out["losses"], "violating rows %.2f" % float((prev_scores < 0).float().mean()),
          "|g_w0| %.3g |g_w2| %.3g |g_w4| %.3g" % tuple(float(np.abs(out["g_w%d" % li]).max()) for li in (0, 2, 4)))


DDPM_FLAGS = ["-e", "e5_ddpm", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--skip_nusc_load"]  # README step 1


def gen_ddpm_step(bs=3, S=16, seed=2006):
    """One denoiser training step of the reference (README step 1; nusc_train.py:539-555 diffusion_prep, :1352-1356
    net(...), :436 loss_diffusion, Adam over net.parameters()): the drawn noise / timesteps / noised commands, eps,
    the loss, the gradients of policy_net (full) and of the encoders (first / last layers full, middle layer by its
    norm and a 16 x 16 corner)."""
    T, args = ref_shim.load(DDPM_FLAGS + ["--n_randoms", str(S)])
    import nusc_model
    nt = args.nt
    N = bs * S * 3
    batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
    b = {k: v.clone() for k, v in batch.items()}
    b["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
    b = T.augment_batch_data(b, b["pre_stlp"].reshape(bs, S, 3, 6)[:, 0, 0], args)
    net = nusc_model.Net(args)
    sd = {k: v for k, v in synthetic.make_weights(seed=1007, nt=nt).items() if not k.startswith(("rect_net", "merge_net"))}
    net.load_state_dict(sd, strict=True)
    coeffs = T.get_diffusion_coeffs(args)
    torch.manual_seed(seed)
    noise, steps, _, noised = T.diffusion_prep(b["params"], n_randoms=S, coeffs=coeffs)
    out = {"noise": noise.numpy(), "steps": steps.numpy(), "noised": noised.numpy()}
    est, feature = net(b, ext={"timestep": steps, "highlevel": b["highlevel_dense"], "noise": noised}, get_feature=True)
    est = est.reshape(N, nt * 2)
    # the reference's own loss of this stage (nusc_train.py:435-437 with dense_scores = tj_scores_prior, :1282-1283;
    # the parser forces stl_bc_mask, :1781)
    assert args.stl_bc_mask
    dense_scores = b["tj_scores_prior"].reshape(bs * S, 3)
    dense_valids = b["valids_dense"]
    stl_acc_mask = (dense_scores * dense_valids > 0).float().reshape(bs * S * 3, 1, 1)
    loss = T.mask_mean(torch.square(noise - est), stl_acc_mask.squeeze(-1))
    out["loss_unmasked"] = np.array(torch.mean(torch.square(noise - est)).item())
    out["mask"] = stl_acc_mask.reshape(-1).numpy()
    opt = torch.optim.Adam(net.parameters(), lr=args.lr)
    opt.zero_grad()
    loss.backward()
    out["eps"] = est.detach().numpy()
    out["feature_scene"] = feature.detach().reshape(bs, S * 3, -1)[:, 0].numpy()
    out["loss"] = np.array(loss.item())
    for name, p in net.named_parameters():
        g = p.grad.numpy()
        if name.startswith("policy_net") or ".2." not in name:
            out["g|" + name] = g.copy()
        else:
            out["gn|" + name] = np.array(np.linalg.norm(g))
            out["gc|" + name] = g[:16, :16].copy() if g.ndim == 2 else g[:16].copy()
    np.savez_compressed(os.path.join(HERE, "ddpm_step.npz"), **out)
    print("ddpm_step:", len(out), "arrays; loss %.5f" % loss.item(),
          "|g policy.0| %.3g |g ego_encoder.0| %.3g" % (np.abs(out["g|policy_net.0.weight"]).max(),
                                                         np.abs(out["g|ego_encoder.0.weight"]).max()))


def metric_inputs(bs=5, m=16, nt=20, seed=2004):
    """trajectories / scores / validity for the diversity metrics: rollouts of the synthetic parameter bank, a random
    accept pattern that includes a lane with nothing accepted, one with two samples and one with collinear samples"""
    b = synthetic.make_scene_batch(bs, nt=nt, n_randoms=m, seed=seed)
    g = torch.Generator().manual_seed(seed)
    s0 = b["ego_traj"][:, 0, :4].unsqueeze(1).unsqueeze(1).repeat(1, m, 3, 1)
    u = b["params"]
    tr = [s0]
    for t in range(nt):
        c = tr[-1]
        ds = torch.stack([c[..., 3] * torch.cos(c[..., 2]), c[..., 3] * torch.sin(c[..., 2]), u[..., t, 0], u[..., t, 1]], -1)
        tr.append(c + ds * 0.5)
    trajs = torch.stack(tr, -2)                                  # (bs, m, 3, nt+1, 4)
    scores = torch.rand(bs, m, 3, generator=g) - 0.4
    scores[0, :, 1] = -1.0                                        # nothing accepted
    scores[1, :, 2] = -1.0
    scores[1, :2, 2] = 1.0                                        # two samples: no hull
    trajs[2, :, 0, :, 1] = trajs[2, :, 0, :, 0] * 0.5 + 1.0       # collinear positions at every step
    valids = torch.cat([b["curr_id"], b["left_id"], b["right_id"]], -1).unsqueeze(1).repeat(1, m, 1)
    return b, trajs, scores, valids, u


def gen_metrics(bs=5, m=16, nt=20, seed=2004):
    """nusc_api.measure_diversity / measure_extra_diversity, utils.compute_entropy and nusc_train.compute_ade_fde of the
    unmodified reference on synthetic trajectories"""
    T, args = ref_shim.load(ref_shim.OURS_FLAGS + ["--n_randoms", str(m), "--sampling_size", str(m)])
    b, trajs, scores, valids, u = metric_inputs(bs, m, nt, seed)
    out = {"in_checksum": np.array([checksum(trajs), checksum(scores), checksum(valids)])}
    xy = trajs[..., :-1, :2].reshape(bs, m, 3, nt * 2)
    r = T.napi.measure_diversity(xy, scores, valids, nt)
    out["ma_std"], out["ma_vol"] = np.array(r[0]), np.array(r[1])
    for i in range(4):
        out["std_list|%d" % i] = np.asarray(r[2][i], dtype=np.float64)
        out["vol_list|%d" % i] = np.asarray(r[3][i], dtype=np.float64)
    ex = T.napi.measure_extra_diversity(trajs[..., :-1, :].reshape(bs, m, 3, nt * 4), scores, valids, nt,
                                        u.reshape(bs, m, 3, nt * 2), -args.mul_w_max, args.mul_w_max, -args.mul_a_max,
                                        args.mul_a_max)
    for k, v in ex.items():
        out["extra|" + k] = np.array(float(v))
    ade, fde = T.compute_ade_fde(b["ego_traj"][..., :4], trajs[..., :-1, :4], valids)
    out["ade_fde"] = np.array([float(ade), float(fde)])
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("metrics:", {k: (v.tolist() if v.size < 4 else v.shape) for k, v in out.items()})


def gen2_flags(n=96, nt=20, knei=8, seed=1010):
    """compute_stl_dense / prep_stl_cache of the reference under the flags that change the predicate math
    (SURVEY 8(b)): --inline (lane end-caps, rows placed before / beyond the polylines), --clip_dist, --norm_stl,
    --refined_nL / --refined_nW (anchor grid), --collision_loss (extra signals + the loss term of nusc_train.py:416-420)."""
    T, args = ref_shim.load(ref_shim.OURS_FLAGS)
    out = {}
    variants = {"inline": dict(inline=True), "inline_clip": dict(inline=True, clip_dist=True), "norm": dict(norm_stl=True),
                "nl3w2": dict(refined_nL=3, refined_nW=2), "nl6": dict(refined_nL=6), "coll": dict(collision_loss=1.0)}
    base = {k: getattr(args, k) for k in ("inline", "clip_dist", "norm_stl", "refined_nL", "refined_nW", "collision_loss")}
    for tag, over in variants.items():
        for k, v in base.items():
            setattr(args, k, v)
        for k, v in over.items():
            setattr(args, k, v)
        stls = T.build_stl_cache(args)
        x, idx, mask = synthetic.make_dense_stl_input(n, nt=nt, n_neighbors=knei, seed=seed, endcaps=True, overlap=True)
        out[tag + "|in_checksum"] = np.array([checksum(x[k]) for k in sorted(x)])
        x["ego_traj"] = x["ego_traj"].clone().requires_grad_()
        scores_list, scores, acc, xo = T.compute_stl_dense(x, stls, idx, mask, args, debug=True)
        for k in ("x2curr_d", "x2left_d", "x2right_d", "min_nei_d"):
            out[tag + "|" + k] = xo[k].detach().numpy()
        out[tag + "|scores"] = scores.detach().numpy()
        loss = T.mask_mean(torch.relu(args.stl_nn_thres - scores), mask)
        if args.collision_loss is not None:
            out[tag + "|min_centroid_d"] = xo["min_centroid_d"].detach().numpy()
            out[tag + "|radius_sum"] = xo["radius_sum"].detach().numpy()
            coll_dist = torch.nn.ReLU()(1 - xo["min_centroid_d"] / torch.clip(xo["radius_sum"], 1e-1))
            coll = torch.mean(torch.clip(torch.sum(coll_dist, dim=-1), max=1)) * args.collision_loss
            out[tag + "|loss_coll"] = np.array(coll.item())
            loss = loss + coll
        (g,) = torch.autograd.grad(loss, [x["ego_traj"]])
        out[tag + "|grad_ego"] = g.numpy()
        if tag == "inline":  # how many poses actually sit on an end-cap
            out[tag + "|n_endcap"] = np.array(int((xo["x2curr_d"].detach() != T.napi.compute_t2l_dist(
                x["ego_traj"].detach()[..., :3], x["currlane_wpts"], False, with_angle=True, inline=False)[0]).sum()))
    for k, v in base.items():
        setattr(args, k, v)
    # get_neighbor_trajs (nusc_train.py:51-60)
    b = synthetic.make_scene_batch(3, nt=nt, n_randoms=4, seed=seed + 1)
    out["nei|short"] = T.get_neighbor_trajs(b["neighbors"], nt, args.dt).numpy()
    out["nei|full"] = T.get_neighbor_trajs(b["neighbors"], nt, args.dt, full=True).numpy()
    # the closed-loop pick (nusc_sim.py:677-683), executed from the reference's own source lines
    import textwrap
    src = open(os.path.join(ref_shim.REF_DIR, "nusc_sim.py")).read().split("\n")[676:683]
    assert "scores_all.reshape(n//3, 3)" in src[0] and "sim_traj = ego_trajs[total_idx]" in src[-1], src
    g = torch.Generator().manual_seed(seed + 2)
    n_rows = 192
    env = {"torch": torch, "time": __import__("time"), "n": n_rows, "scores_all": torch.randn(n_rows, generator=g),
           "ego_controls": torch.randn(n_rows, nt, 2, generator=g), "ego_trajs": torch.randn(n_rows, nt + 1, 4, generator=g)}
    env["scores_all"][30] = env["scores_all"][60] = env["scores_all"].max() + 1.0  # a tie between two mode-0 rows
    out["pick|scores"] = env["scores_all"].clone().numpy()
    out["pick|controls"], out["pick|trajs"] = env["ego_controls"].numpy(), env["ego_trajs"].numpy()
    exec(textwrap.dedent("\n".join(src)), env)
    out["pick|idx"] = np.array(int(env["total_idx"]))
    out["pick|highest"] = np.array(float(env["highest_score"]))
    out["pick|ctrl"], out["pick|traj"] = env["sim_ctrl"].numpy(), env["sim_traj"].numpy()
    np.savez_compressed(os.path.join(HERE, "flags.npz"), **out)
    print("flags:", len(out), "arrays; end-cap poses", int(out["inline|n_endcap"]), "loss_coll", float(out["coll|loss_coll"]))


def gen2_sampler_modes():
    """The sampler's other trigger / call modes, from the reference's own run_sampling_test / diffusion_rollout:
    --guidance_freq, --guidance_sets with --guidance_reverse (nusc_train.py:589-598), --refinement (:1034-1071), and a
    mono rollout under --gt_data_training (:570-572)."""
    out = {}
    gflags = [f for f in ref_shim.GUIDE_FLAGS]
    run_pipeline(gflags + ["--guidance_freq", "7"], "freq", bs=1, S=16, seed=2011, out=out)
    run_pipeline(gflags + ["--guidance_sets", "3", "40", "41", "--guidance_reverse"], "sets", bs=1, S=16, seed=2012, out=out)
    run_pipeline(ref_shim.OURS_FLAGS + ["--refinement"], "refine", bs=2, S=16, seed=2013, out=out)
    for k in [k for k in out if k.split("|")[1] in ("feature", "iter_mid", "tj_scores", "rect_first")]:
        del out[k]
    # mono
    S, bs, seed = 16, 2, 2014
    T, args = ref_shim.load(ref_shim.OURS_FLAGS + ["--n_randoms", str(S), "--gt_data_training"])
    import nusc_model
    nt = args.nt
    batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
    net = nusc_model.Net(args)
    net.load_state_dict(synthetic.make_weights(seed=1007, nt=nt), strict=True)
    n = bs * S
    stream = synthetic.noise_stream(seed + 77, n, nt * 2, args.diffusion_steps - 1)
    it = iter(stream)
    real = torch.randn_like
    torch.randn_like = lambda t, **k: next(it).to(t.dtype)
    try:
        with torch.no_grad():
            feature = net.encode_feat(batch)
            gt_stlp = batch["pre_stlp"].reshape(bs, S, 3, 6)[:, 0, 0]
            hl = batch["gt_high_level"]
            res = T.diffusion_rollout(torch.zeros(n, nt * 2), net, batch, hl, feature, args, T.get_diffusion_coeffs(args),
                                      mono=True, tmp_stlp=gt_stlp)
    finally:
        torch.randn_like = real
    out["mono|final"] = res[0].numpy()
    out["mono|iter_mid"] = res[1][50].numpy()
    np.savez_compressed(os.path.join(HERE, "sampler_modes.npz"), **out)
    print("sampler_modes:", sorted(out))


def gen2_rect_variants(bs=2, S=16, seed=2015):
    """Net.rect_forward of the reference for the RefineNet variants (nusc_model.py:182-235): --diverse_fuse_type cat,
    --no_arch, no --diverse_loss."""
    out = {}
    base = ["-e", "x", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--rect_head", "--flex", "--skip_nusc_load",
            "--n_randoms", str(S)]
    for tag, extra, ex_in in (("cat", ["--diverse_loss", "--diverse_fuse_type", "cat"], 40),
                              ("noarch", ["--diverse_loss", "--no_arch"], 0), ("plain", [], 0), ("clip", ["--diverse_loss", "--clip_rect"], 0)):
        T, args = ref_shim.load(base + extra)
        import nusc_model
        nt = args.nt
        net = nusc_model.Net(args)
        sd = synthetic.make_weights(seed=1007, nt=nt, rect_extra_in=ex_in)
        if not args.diverse_loss:
            sd = {k: v for k, v in sd.items() if not k.startswith("merge_net")}
        net.load_state_dict(sd, strict=True)
        n = bs * S * 3
        g = torch.Generator().manual_seed(seed)
        batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
        with torch.no_grad():
            feat = net.encode_feat(batch).reshape(bs, 1, -1).repeat(1, S * 3, 1).reshape(n, -1)
            hl = torch.tensor([0.0, 1.0, 2.0]).repeat(bs * S).reshape(n, 1)
            stlp = batch["pre_stlp"].reshape(n, 6)
            u0 = torch.stack([(torch.rand(n, nt, generator=g) - 0.5) * 0.8, (torch.rand(n, nt, generator=g) - 0.5) * 8], -1)
            sc = torch.randn(n, generator=g)
            out[tag + "|out"] = net.rect_forward(feat, hl, stlp, u0, sc).numpy()
        out[tag + "|u0"], out[tag + "|scores"] = u0.numpy(), sc.numpy()
    np.savez_compressed(os.path.join(HERE, "rect_variants.npz"), **out)
    print("rect_variants:", sorted(out))


def main():
    """``python tests/golden/make_golden.py [name ...]``: all fixtures, or only the named generators."""
    torch.set_num_threads(8)
    T, args = ref_shim.load(ref_shim.OURS_FLAGS)
    gens = [("stl_kats", gen_stl_kats), ("dense", lambda: gen_dense(T, args)), ("pipeline", gen_pipeline),
            ("trajopt", gen_trajopt), ("metrics", gen_metrics), ("losses", gen_losses), ("refine_step", gen_refine_step),
            ("ddpm_step", gen_ddpm_step)]
    gens += [(k[5:], v) for k, v in sorted(globals().items()) if k.startswith("gen2_")]
    want = sys.argv[1:]
    unknown = [w for w in want if w not in dict(gens)]
    assert not unknown, "unknown generators %s; have %s" % (unknown, [g[0] for g in gens])
    for name, fn in gens:
        if not want or name in want:
            fn()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
