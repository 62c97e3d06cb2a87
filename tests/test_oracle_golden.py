"""Pin the oracle (oracle/pstl_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import pstl_b200  # noqa: F401
from pstl_b200 import synthetic
from oracle import pstl_oracle as O
from formulas import recipes, TupleNS
from make_golden import kat_inputs, checksum

RTOL = 1e-5


def close(a, b, rtol=RTOL, atol=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    assert (np.isfinite(a) == fin).all()
    assert (a[~fin] == b[~fin]).all() or (np.isnan(a[~fin]) == np.isnan(b[~fin])).all()
    if atol is None:
        atol = rtol * max(1.0, float(np.abs(b[fin]).max()) if fin.any() else 1.0)
    np.testing.assert_allclose(a[fin], b[fin], rtol=rtol, atol=atol)


def test_stl_kats(golden_dir):
    G = np.load(os.path.join(golden_dir, "stl_kats.npz"))
    x0 = kat_inputs()
    assert np.allclose([checksum(x0[k]) for k in "abc"], G["in_checksum"])
    fs = recipes(TupleNS)
    for name in G["names"]:
        name = str(name)
        for tau in (1.0, 100.0):
            for hard in (False, True):
                key = "%s|%g|%d" % (name, tau, int(hard))
                x = {k: v.clone().requires_grad_() for k, v in x0.items()}
                y = O.stl_eval(fs[name], x, tau, hard)
                close(y.detach().numpy(), G[key])
                if key + "|ga" in G.files:
                    grs = torch.autograd.grad(y[:, 0].sum(), [x["a"], x["b"], x["c"]], allow_unused=True)
                    for k, g in zip("abc", grs):
                        g = torch.zeros_like(x0[k]) if g is None else g
                        close(g.numpy(), G[key + "|g" + k], atol=1e-6)


def test_seed_kats_from_survey(golden_dir):
    G = np.load(os.path.join(golden_dir, "stl_kats.npz"))
    # SURVEY.md §8(c) table values (captured independently from the reference)
    np.testing.assert_allclose(G["seed|always_0_3"][0], [-0.2, -0.2, -0.4, -0.4, -0.4, 0.05, 0.05, 0.6], atol=2e-6)
    ev = G["seed|eventually_1_4"][0]
    np.testing.assert_allclose(ev[:7], [0.5, 0.5, 0.25, 0.25, 0.6, 0.6, 0.6], atol=2e-6)
    assert ev[7] == -np.inf
    np.testing.assert_allclose(G["seed|ev_alw_and|tau1"][0],
                               [-1.0503844, -0.8605880, -0.5966869, -0.2603241, 0.4093887, 0.3100896, 0.1453476,
                                -0.1709571], atol=2e-6)
    assert str(G["str_symbol"]) == "♢[0:5] (◻[0:9] ((a) & (b)))"
    assert str(G["str_word"]) == "EVENTUALLY[0:5] (ALWAYS[0:9] ((a) AND (b)))"


@pytest.mark.parametrize("tag,n,nt,knei,seed", [("t20k8", 192, 20, 8, 1008), ("t50k16", 48, 50, 16, 1009)])
def test_dense_scores(golden_dir, tag, n, nt, knei, seed):
    G = np.load(os.path.join(golden_dir, "stl_dense.npz"))
    x, idx, mask = synthetic.make_dense_stl_input(n, nt=nt, n_neighbors=knei, seed=seed)
    assert np.allclose([checksum(x[k]) for k in sorted(x)], G[tag + "|in_checksum"])
    x["ego_traj"] = x["ego_traj"].clone().requires_grad_()
    sc = O.stl_scores(x, idx[:, 0], 100.0)
    for k in ("x2curr_d", "x2curr_th", "x2left_d", "x2left_th", "x2right_d", "x2right_th", "min_nei_d"):
        close(x[k].detach().numpy(), G[tag + "|" + k])
    close(sc.detach().numpy(), G[tag + "|scores"])
    loss = O.mask_mean(torch.relu(0.0005 - sc), mask)
    (g,) = torch.autograd.grad(loss, [x["ego_traj"]])
    close(g.numpy(), G[tag + "|grad_ego"], atol=1e-9)


def _pipeline(tag, flags_guidance, K, n_rolls, seed):
    bs, S, nt = 2, 16, 20
    b = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
    W = synthetic.make_weights(1007, nt=nt)
    N = bs * S * 3
    stream = synthetic.noise_stream(seed + 77, N, nt * 2, 99)
    g = dict(before=10, lr=0.01, thres=0.0005, niters=1) if flags_guidance else None
    return O.pipeline(W, b, stream[0], stream[1:], S=S, K=K, n_rolls=n_rolls, guidance=g, n_randoms=S), b, stream, W


def test_pipeline_ours(golden_dir):
    G = np.load(os.path.join(golden_dir, "pipeline.npz"))
    out, b, stream, W = _pipeline("ours", False, 5, 0, 2001)
    want = G["ours|in_checksum"]
    got = [checksum(b[k]) for k in sorted(b)] + [checksum(stream[0]), sum(checksum(v) for v in W.values())]
    assert np.allclose(got, want)
    close(out["feature"].numpy(), G["ours|feature"])
    close(out["final_iterate"].numpy(), G["ours|final_iterate"])
    close(out["cand_scores"].numpy(), G["ours|cand_scores"])
    close(out["controls"].numpy(), G["ours|controls"])
    close(out["scores"].numpy(), G["ours|scores"])


def test_pipeline_guidance(golden_dir):
    G = np.load(os.path.join(golden_dir, "pipeline.npz"))
    out, *_ = _pipeline("guide", True, 10, 3, 2002)
    assert int(G["guide|n_guidance_calls"]) == 10
    close(out["final_iterate"].numpy(), G["guide|final_iterate"])
    close(out["cand_scores"].numpy(), G["guide|cand_scores"])
    close(out["controls"].numpy(), G["guide|controls"])
    close(out["scores"].numpy(), G["guide|scores"])


def test_trajopt_loop(golden_dir):
    """oracle trajopt == the reference's trajectory-optimisation loop (nusc_train.py:1303-1325) on one synthetic batch"""
    G = np.load(os.path.join(golden_dir, "trajopt.npz"))
    lr, thres, reg, w_max, a_max, iters = [float(v) for v in G["hyper"]]
    iters = int(iters)
    bs, S_, nt = 2, 16, 20
    b = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S_, seed=2003)
    np.testing.assert_allclose(np.array([checksum(b[k]) for k in sorted(b)]), G["in_checksum"], rtol=1e-12)
    rec = {}

    def record(ii, loss, dl, rl, scores, g, p):
        rec["loss|%d" % ii] = np.array([loss, dl, rl])
        rec["scores|%d" % ii] = scores.numpy()
        rec["grad|%d" % ii] = g.numpy()
        rec["params|%d" % ii] = p.numpy()

    O.trajopt(b, S_, nt, 0.5, iters, lr=lr, thres=thres, reg=reg, w_max=w_max, a_max=a_max, record=record)
    for ii in range(iters):
        np.testing.assert_allclose(rec["loss|%d" % ii], G["loss|%d" % ii], rtol=2e-5, atol=1e-7)
    close(rec["scores|0"], G["scores|0"])
    close(rec["grad|0"], G["grad|0"], rtol=2e-4)
    close(rec["params|0"], G["params|0"])
    # Adam turns rounding-level gradient differences of the |g| ~ 1e-8 elements into lr-sized moves (as in guidance):
    # later iterates are compared on the bulk
    for ii in (4, iters - 1):
        err = np.abs(rec["params|%d" % ii] - G["params|%d" % ii])
        assert np.percentile(err, 99) < 1e-5 and err.max() < 2 * lr * (ii + 1), (ii, np.percentile(err, 99), err.max())
    err = np.abs(rec["scores|%d" % (iters - 1)] - G["scores|%d" % (iters - 1)])
    assert np.percentile(err, 99) < 1e-4 * max(1.0, np.abs(G["scores|%d" % (iters - 1)]).max())



def test_diversity_metrics(golden_dir):
    """oracle diversity (masked std + scipy hull areas) and the device-agnostic metric functions of pstl_b200.metrics
    against nusc_api.measure_diversity / measure_extra_diversity / compute_ade_fde of the reference"""
    from make_golden import metric_inputs
    from pstl_b200 import metrics as M
    G = np.load(os.path.join(golden_dir, "metrics.npz"))
    bs, m, nt = 5, 16, 20
    b, trajs, scores, valids, u = metric_inputs(bs, m, nt, 2004)
    np.testing.assert_allclose([checksum(trajs), checksum(scores), checksum(valids)], G["in_checksum"], rtol=1e-12)
    std, vol, ma_std, ma_vol = O.diversity(trajs[..., :-1, :2].reshape(bs, m, 3, nt * 2), scores, valids, nt)
    np.testing.assert_allclose(ma_std, G["ma_std"], rtol=1e-5)
    np.testing.assert_allclose(ma_vol, G["ma_vol"], rtol=1e-6)
    val = valids[:, 0, :].numpy() != 0
    for i in range(3):
        np.testing.assert_allclose(std[:, i] * val[:, i], G["std_list|%d" % (i + 1)], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(vol[:, i] * val[:, i], G["vol_list|%d" % (i + 1)], rtol=1e-6, atol=1e-9)
    ex = M.measure_extra_diversity(trajs[..., :-1, :].reshape(bs, m, 3, nt * 4), scores, valids, nt,
                                   u.reshape(bs, m, 3, nt * 2), -0.5, 0.5, -5.0, 5.0)
    for k in ("ent_s", "ent_w", "ent_a", "ent_wa", "area"):
        np.testing.assert_allclose(float(ex[k]), G["extra|" + k], rtol=1e-5)
    ade, fde = M.compute_ade_fde(b["ego_traj"][..., :4], trajs[..., :-1, :4], valids)
    np.testing.assert_allclose([float(ade), float(fde)], G["ade_fde"], rtol=1e-6)


LOSS_TAGS = ("ours", "weighted", "detach", "plain")
LOSS_KEYS = ("loss", "loss_stl", "loss_reg", "loss_diversity", "extra_loss_reg")


def loss_kwargs(G, tag):
    thres, stl_w, dscale, dweight, reg_w, extra_w, n_shards, diverse, detach, w_max, a_max = [float(v) for v in G[tag + "|hyper"]]
    bs, S_, nt = [int(v) for v in G["shape"]]
    return dict(n_scenes=bs, S=S_, nt=nt, n_shards=int(n_shards), diverse_loss=bool(diverse), diverse_detach=bool(detach),
                w_max=w_max, a_max=a_max, stl_nn_thres=thres, stl_weight=stl_w, diversity_scale=dscale,
                diversity_weight=dweight, rect_reg_loss=reg_w, extra_rect_reg=extra_w)


@pytest.mark.parametrize("tag", LOSS_TAGS)
def test_refine_losses(golden_dir, tag):
    """oracle refine_losses == the reference's compute_policy_loss (nusc_train.py:370-478, --rect_head training step):
    loss terms, d loss / d rect_controls with the scores held fixed, and d loss / d scores"""
    G = np.load(os.path.join(golden_dir, "losses.npz"))
    kw = loss_kwargs(G, tag)
    rect = torch.from_numpy(G["rect_controls"]).requires_grad_()
    scores = torch.from_numpy(G[tag + "|scores"]).requires_grad_()
    out = O.refine_losses(rect, torch.from_numpy(G["nn_controls"]), scores, torch.from_numpy(G["valid"]), **kw)
    for i, k in enumerate(LOSS_KEYS):
        if np.isfinite(G[tag + "|losses"][i]):
            np.testing.assert_allclose(float(out[k].detach()), G[tag + "|losses"][i], rtol=2e-5, atol=1e-7, err_msg=k)
    g_rect, g_sc = torch.autograd.grad(out["loss"], [rect, scores], allow_unused=True)
    close(g_rect.numpy(), G[tag + "|grad_direct"], rtol=2e-5)
    close((g_sc if g_sc is not None else torch.zeros_like(scores)).numpy(), G[tag + "|grad_scores"], rtol=2e-5)


def test_refine_train_step(golden_dir):
    """oracle refine_train_step == one --rect_head training step of the reference (tests/golden/refine_step.npz):
    RefineNet output, loss, d loss / d rect_controls, the six rect_net gradients, and the weights after Adam"""
    G = np.load(os.path.join(golden_dir, "refine_step.npz"))
    Lz = np.load(os.path.join(golden_dir, "losses.npz"))
    kw = loss_kwargs(Lz, "weighted")
    bs, S_, nt = kw["n_scenes"], kw["S"], kw["nt"]
    b = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S_, seed=2005)
    W = synthetic.make_weights(seed=1007, nt=nt)
    r = O.refine_train_step(W, b, torch.from_numpy(G["feat_scene"]), torch.from_numpy(Lz["nn_controls"]), 0.5, **kw)
    close(r["prev_scores"].numpy(), G["prev_scores"])
    close(r["rect"].numpy(), G["rect"])
    for i, k in enumerate(("loss", "loss_stl", "loss_reg", "loss_diversity")):
        np.testing.assert_allclose(float(r["losses"][k].detach()), G["losses"][i], rtol=5e-5, atol=1e-6, err_msg=k)
    close(r["grad_rect"].numpy(), G["grad_rect"], rtol=1e-4)
    for li in (0, 2, 4):
        close(r["grads"]["rect_net.%d.weight" % li].numpy(), G["g_w%d" % li], rtol=1e-4)
        close(r["grads"]["rect_net.%d.bias" % li].numpy(), G["g_b%d" % li], rtol=1e-4)
    params = [r["W"]["rect_net.%d.%s" % (li, k)] for li in (0, 2, 4) for k in ("weight", "bias")]
    torch.optim.Adam(params, lr=float(G["lr"])).step()
    for li in (0, 2, 4):
        # Adam's first step moves every weight by ~lr * sign(g): compare on the bulk, bound the rest by 2 lr
        for k, key in (("weight", "w%d_after"), ("bias", "b%d_after")):
            err = np.abs(r["W"]["rect_net.%d.%s" % (li, k)].detach().numpy() - G[key % li])
            assert np.percentile(err, 99) < 1e-6 and err.max() <= 2.001 * float(G["lr"]), (li, k, err.max())


def check_ddpm_grads(G, grads, rtol):
    """policy_net and first/last encoder layers in full, the encoders' middle layer by norm and a corner"""
    n = 0
    for key in G.files:
        kind, _, name = key.partition("|")
        if kind == "g":
            close(grads[name], G[key], rtol=rtol)
        elif kind == "gn":
            np.testing.assert_allclose(np.linalg.norm(np.asarray(grads[name], np.float64)), float(G[key]), rtol=10 * rtol)
        elif kind == "gc":
            g = np.asarray(grads[name])
            close(g[:16, :16] if g.ndim == 2 else g[:16], G[key], rtol=rtol,
                  atol=rtol * float(np.abs(g).max()))
        else:
            continue
        n += 1
    assert n == 30  # 24 tensors, the three middle encoder layers (weight, bias) counted twice


def test_ddpm_train_step(golden_dir):
    """oracle ddpm_train_step == one denoiser training step of the reference (tests/golden/ddpm_step.npz)"""
    G = np.load(os.path.join(golden_dir, "ddpm_step.npz"))
    bs, S_, nt = 3, 16, 20
    b = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S_, seed=2006)
    r = O.ddpm_train_step(synthetic.make_weights(1007, nt=nt), b, torch.from_numpy(G["noise"]),
                          torch.from_numpy(G["steps"]), torch.from_numpy(G["noised"]), S_, nt)
    close(r["feature"].numpy(), G["feature_scene"])
    close(r["eps"].numpy(), G["eps"])
    np.testing.assert_allclose(float(r["loss"]), float(G["loss"]), rtol=1e-5)
    check_ddpm_grads(G, {k: v.numpy() for k, v in r["grads"].items()}, 1e-4)
