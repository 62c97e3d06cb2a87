"""ncu target: a few launches of the dense time-parallel scorer at config 5's first cell"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pstl_b200
from pstl_b200 import nusc_train as NT
from bench import _dense_rows_on_device
n, T, K = int(sys.argv[1]) if len(sys.argv) > 1 else 262144, int(sys.argv[2]) if len(sys.argv) > 2 else 20, int(sys.argv[3]) if len(sys.argv) > 3 else 8
x, idx, mask = _dense_rows_on_device(n, T, K, torch.device("cuda"), 7)
args = NT.default_args(nt=T)
stls = NT.build_stl_cache(args)
for _ in range(4):
    NT.compute_stl_dense(x, stls, idx, mask, args)
torch.cuda.synchronize()
