"""NCCL check of sharding.gather_scores on real GPUs (the pytest suite covers the gloo path on CPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_gather_check.py
"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pstl_b200
from pstl_b200 import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
sc = (torch.arange(1000, dtype=torch.float32) + 1000 * rank).cuda()
ix = (torch.arange(1000, dtype=torch.int32) * 3 + rank).cuda()
a, b = sharding.gather_scores(sc, ix, equal_sizes=True)
ref_a = torch.cat([torch.arange(1000, dtype=torch.float32) + 1000 * r for r in range(world)]).cuda()
ref_b = torch.cat([torch.arange(1000, dtype=torch.int32) * 3 + r for r in range(world)]).cuda()
assert torch.equal(a, ref_a) and torch.equal(b, ref_b) and b.dtype == torch.int32
a2, b2 = sharding.gather_scores(sc, ix)
assert torch.equal(a2, ref_a) and torch.equal(b2, ref_b)
# the side-stream gather bench.py uses: results equal the blocking one, also when the inputs are overwritten right after
g = sharding.AsyncScoreGather(1000, torch.device("cuda", int(os.environ["LOCAL_RANK"])))
for rep in range(3):
    s2, i2 = sc + rep, ix + rep
    g.submit(s2, i2)
    s2.add_(1000.0)  # the next batch overwrites the tensors: submit() has already snapshotted them
    a3, b3 = g.result()
    torch.cuda.synchronize()
    assert torch.equal(a3, ref_a + rep) and torch.equal(b3, ref_b + rep), rep
    assert g.last_us() is not None and g.last_us() > 0
if rank == 0: print("gather ok")
dist.destroy_process_group()
