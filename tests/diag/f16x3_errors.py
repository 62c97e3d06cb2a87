"""Where the split-operand (f16x3) path's error sits: per-stage worst deviation from the CPU oracle on the smoke() batch and
a larger one, next to the fp32 SIMT path.   python tests/diag/f16x3_errors.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pstl_b200
from pstl_b200 import synthetic
from pstl_b200.nusc_train import build_stl_cache, default_args, get_diffusion_coeffs, sample_and_score
from pstl_b200.nusc_model import Net
from oracle import pstl_oracle as O

for bs, S, seed in ((2, 16, 2001), (6, 32, 2005)):
    nt = 20
    args = default_args(n_randoms=S, sampling_size=S)
    batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=seed)
    W = synthetic.make_weights(1007, nt=nt)
    N = bs * S * 3
    stream = synthetic.noise_stream(seed + 77, N, nt * 2, 99)
    net = Net(args); net.load_state_dict(W, strict=True); net = net.cuda()
    args.inject_noise = [t.cuda() for t in stream]
    ref = O.pipeline(W, batch, stream[0], stream[1:], S=S, K=5, n_rolls=0, n_randoms=S)
    bc = {k: v.cuda() for k, v in batch.items()}
    sc = {"final_iterate": np.array([0.5, 5.0]), "cand_scores": 1.0, "controls": np.array([0.5, 5.0]), "scores": 1.0}
    for prec in ("fp32", "f16x3"):
        args.precision = prec
        out = sample_and_score(net, bc, build_stl_cache(args), get_diffusion_coeffs(args), args)
        torch.cuda.synchronize()
        msg = []
        for k, s in sc.items():
            a, b = out[k].cpu().numpy(), ref[k].numpy()
            e = np.abs(a - b) / np.maximum(s, np.abs(b))
            msg.append("%s max %.2e p99 %.2e" % (k, e.max(), np.percentile(e, 99)))
        same = (out["best_idx"].cpu().numpy() == ref["best_idx"].numpy()).mean()
        print("N=%d %-7s %s | same candidate %.3f" % (N, prec, " | ".join(msg), same))
