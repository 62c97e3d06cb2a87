"""Pin the oracle's flag variants against the unmodified reference (tests/golden/flags.npz, sampler_modes.npz; generators
gen2_* in tests/golden/make_golden.py): --inline / --clip_dist / --norm_stl / --refined_nL,nW / --collision_loss in the
predicates, the guidance triggers --guidance_freq / --guidance_sets / --guidance_reverse, --refinement.  CPU only."""
import os

import numpy as np
import pytest
import torch

import pstl_b200  # noqa: F401
from pstl_b200 import synthetic
from oracle import pstl_oracle as O
from make_golden import checksum
from test_oracle_golden import close

VARIANTS = {"inline": dict(inline=True), "inline_clip": dict(inline=True, clip_dist=True), "norm": dict(norm=True),
            "nl3w2": dict(nL=3, nW=2), "nl6": dict(nL=6), "coll": dict(collision=True)}


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_flag_variants(golden_dir, tag):
    G = np.load(os.path.join(golden_dir, "flags.npz"))
    x, idx, mask = synthetic.make_dense_stl_input(96, nt=20, n_neighbors=8, seed=1010, endcaps=True, overlap=True)
    assert np.allclose([checksum(x[k]) for k in sorted(x)], G[tag + "|in_checksum"])
    x["ego_traj"] = x["ego_traj"].clone().requires_grad_()
    sc = O.stl_scores(x, idx[:, 0], 100.0, **VARIANTS[tag])
    for k in ("x2curr_d", "x2left_d", "x2right_d", "min_nei_d"):
        close(x[k].detach().numpy(), G[tag + "|" + k])
    close(sc.detach().numpy(), G[tag + "|scores"])
    loss = O.mask_mean(torch.relu(0.0005 - sc), mask)
    if tag == "coll":
        close(x["min_centroid_d"].detach().numpy(), G["coll|min_centroid_d"])
        close(x["radius_sum"].detach().numpy(), G["coll|radius_sum"])
        coll = O.collision_loss(x, 1.0)
        np.testing.assert_allclose(float(coll.detach()), float(G["coll|loss_coll"]), rtol=1e-6)
        assert float(coll) > 0
        loss = loss + coll
    (g,) = torch.autograd.grad(loss, [x["ego_traj"]])
    close(g.numpy(), G[tag + "|grad_ego"], atol=1e-9)
    if tag == "inline":
        assert int(G["inline|n_endcap"]) > 100  # the fixture does exercise the end-cap branch


def _pipeline(seed, bs, K, n_rolls, guidance, **kw):
    S, nt = 16, 20
    b = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
    W = synthetic.make_weights(1007, nt=nt)
    stream = synthetic.noise_stream(seed + 77, bs * S * 3, nt * 2, 99)
    return O.pipeline(W, b, stream[0], stream[1:], S=S, K=K, n_rolls=n_rolls, guidance=guidance, n_randoms=S, **kw)


@pytest.mark.parametrize("tag,seed,trig,n_calls", [("freq", 2011, dict(freq=7), 14),
                                                   ("sets", 2012, dict(sets=[3, 40, 41], reverse=True), 3)])
def test_guidance_triggers(golden_dir, tag, seed, trig, n_calls):
    G = np.load(os.path.join(golden_dir, "sampler_modes.npz"))
    assert int(G[tag + "|n_guidance_calls"]) == n_calls == len(O.guidance_steps(100, **trig))
    out = _pipeline(seed, 1, 10, 3, dict(lr=0.01, thres=0.0005, niters=1, **trig))
    close(out["final_iterate"].numpy(), G[tag + "|final_iterate"])
    close(out["cand_scores"].numpy(), G[tag + "|cand_scores"])
    close(out["controls"].numpy(), G[tag + "|controls"])
    close(out["scores"].numpy(), G[tag + "|scores"])


def test_refinement(golden_dir):
    G = np.load(os.path.join(golden_dir, "sampler_modes.npz"))
    out = _pipeline(2013, 2, 5, 0, None, refine_mix=True)
    close(out["rect_controls"].numpy(), G["refine|controls"])
    moved = np.abs(G["refine|final_controls"] - G["refine|controls"]).max(axis=(1, 2)) > 1e-6
    assert moved.any()  # the fixture does exercise the mixing branch
    close(out["controls"].numpy(), G["refine|final_controls"], rtol=1e-4)
    close(out["scores"].numpy(), G["refine|scores"], rtol=1e-4)
