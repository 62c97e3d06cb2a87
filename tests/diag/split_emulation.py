"""CPU emulation behind DESIGN.md section 3.6: the oracle's 99-step sampler with the policy_net matmuls replaced by
split-operand products (pieces rounded to bf16 or fp16, fp32 accumulation), against an fp64 run of the same chain.
    python tests/diag/split_emulation.py          # ~1 minute on CPU, no GPU needed
Prints the worst deviation of the final controls in units of the control range for: fp32, two bf16 pieces (3 products),
three bf16 pieces (6 products), two fp16 pieces (3 products)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pstl_b200  # noqa: E402,F401
from pstl_b200 import synthetic  # noqa: E402
from oracle import pstl_oracle as O  # noqa: E402

bs, S, nt = 4, 16, 20
W = synthetic.make_weights(1007, nt=nt)
batch = synthetic.make_scene_batch(bs, nt=nt, n_randoms=S, seed=2001)
N = bs * S * 3
stream = synthetic.noise_stream(2078, N, nt * 2, 99)
dense = O.densify(batch, S, nt)
feat = O.encode_scene(W, batch)
feat_dense = feat.reshape(bs, 1, -1).expand(bs, S * 3, feat.shape[-1]).reshape(N, -1)
hl, stlp = dense["mode"][:, None], dense["stlp"][:, 0]
cfg = {"mode": "fp32", "piece": torch.bfloat16}


def rnd(x):
    return x.to(cfg["piece"]).to(torch.float32)


def lin_split(x, w, b, terms):
    xs, ws, rx, rw = [], [], x, w
    for _ in range(terms):
        p = rnd(rx); xs.append(p); rx = rx - p
        p = rnd(rw); ws.append(p); rw = rw - p
    acc = 0
    for i in range(terms):          # products with i + j < terms: 3 for two pieces, 6 for three
        for j in range(terms - i):
            acc = acc + xs[i] @ ws[j].T
    return acc + b


orig = O.mlp3


def mlp3(Wd, name, x):
    m = cfg["mode"]
    if name != "policy_net" or m == "fp32":
        return orig(Wd, name, x)
    if m == "fp64":
        lin = torch.nn.functional.linear
        h = torch.relu(lin(x.double(), Wd[name + ".0.weight"].double(), Wd[name + ".0.bias"].double()))
        h = torch.relu(lin(h, Wd[name + ".2.weight"].double(), Wd[name + ".2.bias"].double()))
        return lin(h, Wd[name + ".4.weight"].double(), Wd[name + ".4.bias"].double()).float()
    h = torch.relu(lin_split(x, Wd[name + ".0.weight"], Wd[name + ".0.bias"], m))
    h = torch.relu(lin_split(h, Wd[name + ".2.weight"], Wd[name + ".2.bias"], m))
    return lin_split(h, Wd[name + ".4.weight"], Wd[name + ".4.bias"], m)


O.mlp3 = mlp3
runs = [("fp64", "fp64", torch.bfloat16), ("fp32", "fp32", torch.bfloat16), ("two bf16 pieces", 2, torch.bfloat16),
        ("three bf16 pieces", 3, torch.bfloat16), ("two fp16 pieces", 2, torch.float16)]
res = {}
for name, mode, piece in runs:
    cfg["mode"], cfg["piece"] = mode, piece
    res[name] = O.ddpm_sample(W, feat_dense, hl, stlp, stream[0], stream[1:], steps=100, nt=nt, clip=False)[-1]
scale = torch.tensor([0.5, 5.0])
for name, _, _ in runs[1:]:
    e = ((res[name] - res["fp64"]) / scale).abs()
    print("%-18s max deviation from fp64 %.3g   mean %.3g   (units of the control range)" % (name, e.max().item(), e.mean().item()))
