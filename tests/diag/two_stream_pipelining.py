"""Experiment: two CapturedPipeline runners replayed alternately on two CUDA streams (batch i+1's front end under batch i's
tail) against one runner on one stream.   python tests/diag/two_stream_pipelining.py [scenes] [depth]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import pstl_b200
from pstl_b200 import synthetic, nusc_train as NT
from pstl_b200.nusc_model import Net

scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
args = NT.default_args(precision="bf16")
net = Net(args); net.load_state_dict(synthetic.make_weights(1007)); net = net.cuda()
stls, co = NT.build_stl_cache(args), NT.get_diffusion_coeffs(args)
b = {k: v.cuda() for k, v in synthetic.make_scene_batch(scenes, seed=3).items()}
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2
R = [NT.CapturedPipeline(net, stls, co, args, b) for _ in range(D)]
S = [torch.cuda.Stream() for _ in range(D)]
flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
K = 24

def run(two):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if two:
        for s in S: s.wait_stream(torch.cuda.current_stream())
    for i in range(K):
        if two:
            with torch.cuda.stream(S[i % D]):
                flush.fill_(i & 255)
                R[i % D](b)
        else:
            flush.fill_(i & 255)
            R[0](b)
    if two:
        for s in S: torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K

for _ in range(2):
    print("one stream  %.3f ms/step   %d streams %.3f ms/step" % (run(False), D, run(True)))
o0, o1 = R[0](b), R[1](b)
torch.cuda.synchronize()
print("finite:", bool(torch.isfinite(o0["scores"]).all() and torch.isfinite(o1["scores"]).all()))
