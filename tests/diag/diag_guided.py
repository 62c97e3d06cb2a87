"""diagnostic (not a test): error statistics of the guided steps / refinement / bf16 path vs the golden fixtures"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import pstl_b200
from pstl_b200 import nusc_train as NT
from test_gpu_flags import _pipeline, npy
GD = os.path.join(ROOT, "tests", "golden")

def guided(Gf, tag, flags, seed, bs):
    G = np.load(os.path.join(GD, Gf))
    out, net, batch, args = _pipeline(flags, seed, bs)
    beta = NT.get_diffusion_coeffs(args)[0].cpu().numpy()
    pack, stls = out["pack"], NT.build_stl_cache(args)
    mu_in, g_ref, mu_out, sc = (G[tag + "|gstep_" + k] for k in ("mu_in", "grad", "mu_out", "scores"))
    steps = [i for i in range(99, 0, -1) if NT.guidance_step_mask(args)[i]]
    for j, i in enumerate(steps):
        mu = torch.from_numpy(mu_in[j]).cuda().contiguous()
        mu, grad, _ = NT.guidance_step(pack, mu, stls, args, float(beta[i]))
        gr, g0 = npy(grad).reshape(pack.N, -1), g_ref[j]
        scale = np.maximum(np.abs(g0).max(axis=1, keepdims=True), 1e-30)
        rel = np.abs(gr - g0) / scale
        err = np.abs(npy(mu).reshape(pack.N, -1) - mu_out[j])
        bad = np.argwhere(err > 1e-5)
        print(tag, "step", i, "grad rel-to-rowmax: max %.2e p99 %.2e | mu err max %.2e n>1e-5 %d | rows with g_ref==0 row: %d, ours nonzero there: %d"
              % (rel.max(), np.percentile(rel, 99), err.max(), len(bad), (np.abs(g0).max(1) == 0).sum(), ((np.abs(g0).max(1) == 0) & (np.abs(gr).max(1) != 0)).sum()))
        for (r, c) in bad[:6]:
            print("    row %d col %d: g_ref %.3e ours %.3e  mu_ref %.6f ours %.6f score_ref %.6f" % (r, c, g0[r, c], gr[r, c], mu_out[j][r, c], npy(mu).reshape(pack.N, -1)[r, c], sc[j][r]))

guided("pipeline.npz", "guide", NT.GUIDANCE_FLAGS, 2002, 2)
guided("sampler_modes.npz", "sets", NT.GUIDANCE_FLAGS + ["--guidance_sets", "3", "40", "41", "--guidance_reverse"], 2012, 1)

# refinement
G = np.load(os.path.join(GD, "sampler_modes.npz"))
out, net, batch, args = _pipeline(NT.OURS_FLAGS + ["--refinement"], 2013, 2)
a, b, base = npy(out["controls"]), G["refine|final_controls"], G["refine|controls"]
moved = np.abs(b - base).reshape(b.shape[0], -1).max(axis=1) > 1e-6
err = (np.abs(a - b) / np.array([0.5, 5.0])).reshape(a.shape[0], -1).max(axis=1)
print("refine: moved rows", moved.sum(), "err on moved sorted:", np.sort(err[moved])[::-1][:20])
sa, sb = npy(out["scores"]), G["refine|scores"]
print("refine scores err sorted", np.sort(np.abs(sa - sb))[::-1][:12], "acc ours %.3f ref %.3f" % ((sa > 0).mean(), (sb > 0).mean()))

# bf16
G = np.load(os.path.join(GD, "pipeline.npz"))
out, net, batch, args = _pipeline(NT.OURS_FLAGS, 2001, 2, precision="bf16")
cs, cref = npy(out["cand_scores"]), G["ours|cand_scores"]
srt = np.sort(cref, axis=0)
marg = srt[-1] - srt[-2]
e_it = np.abs(npy(out["final_iterate"]) - G["ours|final_iterate"]) / np.array([0.5, 5.0])
print("bf16: iterate err max %.3e p99 %.3e; cand score err max %.3e p99 %.3e p50 %.3e" % (e_it.max(), np.percentile(e_it, 99), np.abs(cs - cref).max(), np.percentile(np.abs(cs - cref), 99), np.median(np.abs(cs - cref))))
print("margins percentiles", np.percentile(marg, [10, 25, 50, 75, 90]))
agree = npy(out["best_idx"]) == cref.argmax(0)
print("best_idx agree %d/%d; disagree margins:" % (agree.sum(), agree.size), np.sort(marg[~agree]))
e_c = (np.abs(npy(out["controls"]) - G["ours|controls"]) / np.array([0.5, 5.0])).reshape(len(agree), -1).max(axis=1)
e_s = np.abs(npy(out["scores"]) - G["ours|scores"])
print("controls err (agreeing rows) max %.3e; scores err max %.3e p99 %.3e" % (e_c[agree].max(), e_s[agree].max(), np.percentile(e_s[agree], 99)))
