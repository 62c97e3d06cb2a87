"""BASELINE config 2 at full size through sample_and_score on the shared Philox stream: the f16x3 and bf16 pipelines against
the fp32 pipeline (final iterate, refined controls, scores, selected candidate).   python tests/diag/full_size_pipeline_errors.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pstl_b200
from pstl_b200 import synthetic, nusc_train as NT
from pstl_b200.nusc_model import Net

bs, S, nt = 1024, 64, 20
W = synthetic.make_weights(1007, nt=nt)
batch = {k: v.cuda() for k, v in synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=3).items()}
outs = {}
for prec in ("fp32", "f16x3", "f16", "bf16"):
    args = NT.default_args(precision=prec)
    args.seed = 99
    net = Net(args); net.load_state_dict(W); net = net.cuda()
    NT._call_counter[0] = 7
    torch.manual_seed(0)
    o = NT.sample_and_score(net, batch, NT.build_stl_cache(args), NT.get_diffusion_coeffs(args), args)
    torch.cuda.synchronize()
    outs[prec] = {k: o[k].float().cpu().numpy() for k in ("final_iterate", "cand_scores", "controls", "scores", "best_idx")}
ref = outs["fp32"]
cs = np.sort(ref["cand_scores"], axis=0)
margin = cs[-1] - cs[-2]
for prec in ("f16x3", "f16", "bf16"):
    o = outs[prec]
    e_it = (np.abs(o["final_iterate"] - ref["final_iterate"]) / np.array([0.5, 5.0])).max()
    e_u = (np.abs(o["controls"] - ref["controls"]) / np.array([0.5, 5.0])).reshape(len(margin), -1).max(1)
    e_s = np.abs(o["scores"] - ref["scores"])
    e_c = np.abs(o["cand_scores"] - ref["cand_scores"]).max(0)
    same = o["best_idx"] == ref["best_idx"]
    print("%-6s iterate max %.2e | cand_scores max %.2e p99.9 %.2e | same candidate %.5f (differing rows: largest margin %.2e) | "
          "controls max %.2e p99.9 %.2e | scores max %.2e p99.9 %.2e (rows with the same candidate: max %.2e)"
          % (prec, e_it, e_c.max(), np.percentile(e_c, 99.9), same.mean(), margin[~same].max() if (~same).any() else 0.0,
             e_u.max(), np.percentile(e_u, 99.9), e_s.max(), np.percentile(e_s, 99.9), e_s[same].max()))
