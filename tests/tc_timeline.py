"""Print the clock64 timeline of one tile-step of k_denoiser_tc (or, with argument 2, of one step of the CTA-pair
engine k_denoiser_tc2).  Needs a developer build of the library:
    PSTL_BUILD_TC_DEBUG=1 python pstl-diffusion-policy_b200/build.py --force && python tests/tc_timeline.py [1|2]
(the product build compiles the instrumentation out; rebuild without the variable afterwards)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PSTL_TC_DEBUG"] = "1"
import torch
import pstl_b200
from pstl_b200 import synthetic, nusc_train as NT
from pstl_b200.nusc_model import Net
args = NT.default_args(precision="bf16", tc_engine=int(sys.argv[1]) if len(sys.argv) > 1 else 1)
net = Net(args); net.load_state_dict(synthetic.make_weights(1007)); net = net.cuda()
b = {k: v.cuda() for k, v in synthetic.make_scene_batch(1024, seed=3).items()}
for _ in range(2):
    NT.sample_and_score(net, b, NT.build_stl_cache(args), NT.get_diffusion_coeffs(args), args)
torch.cuda.synchronize()
