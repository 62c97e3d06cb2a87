"""Scene sharding across the GPUs of one box.

The candidate pipeline has no data-path exchange: every tensor is indexed by scene or by
chain-of-scene, so scenes are block-partitioned over ranks and each rank runs the whole path on
its shard.  The only collective is the gather of per-chain scores / selected indices and a few
metric partial sums at the end of a batch (NCCL over NVLink on GPUs, gloo in the CPU tests).
Guidance couples rows through one scalar (the loss normaliser mean(valid), reference
nusc_train.py:23-27,619): ``guidance_normaliser`` all-reduces it so a batch split over ranks
reproduces the single-batch update exactly.
"""
import torch
import torch.distributed as dist

SCENE_KEYS = ("ego_traj", "neighbors", "neighbors_traj", "currlane_wpts", "leftlane_wpts", "rightlane_wpts",
              "curr_id", "left_id", "right_id", "gt_high_level", "pre_stlp", "params", "params_init",
              "tj_scores_prior", "traj_i", "ti")


def shard_bounds(n_scenes, rank, world):
    """contiguous block partition; the first (n_scenes % world) ranks get one extra scene"""
    base, extra = divmod(n_scenes, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    """slice every per-scene tensor of a batch dict to this rank's scenes"""
    bs = batch["currlane_wpts"].shape[0]
    lo, hi = shard_bounds(bs, rank, world)
    return {k: (v[lo:hi] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == bs else v)
            for k, v in batch.items()}


def guidance_normaliser(valid_local, group=None):
    """(N_total, mean(valid) over all ranks) for the guidance loss; one 2-float all-reduce"""
    part = torch.stack([valid_local.sum(), torch.tensor(float(valid_local.numel()), device=valid_local.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    n_total = float(part[1].item())
    return n_total, float(part[0].item()) / n_total


def gather_scores(scores_local, best_idx_local=None, group=None, equal_sizes=False):
    """all-gather per-chain scores (and int32 selected-candidate indices) in rank order.
    Shards may differ by one scene, so sizes are exchanged first and tensors padded; a caller that knows
    every rank holds the same number of chains passes ``equal_sizes`` and gets one collective per tensor
    with no host read-back."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return scores_local, best_idx_local
    world = dist.get_world_size(group)
    if equal_sizes and scores_local.is_cuda:
        n = scores_local.numel()
        if best_idx_local is not None and best_idx_local.dtype == torch.int32 and scores_local.dtype == torch.float32:
            # ONE collective: the int32 indices travel as raw bits behind the scores
            send = torch.cat([scores_local.reshape(-1), best_idx_local.reshape(-1).view(torch.float32)])
            out = torch.empty(world * 2 * n, dtype=torch.float32, device=send.device)
            dist.all_gather_into_tensor(out, send, group=group)
            out = out.reshape(world, 2, n)
            return out[:, 0].reshape(-1), out[:, 1].contiguous().view(torch.int32).reshape(-1)

        def gather_eq(t):
            out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t.reshape(-1).contiguous(), group=group)
            return out
        return gather_eq(scores_local), (gather_eq(best_idx_local) if best_idx_local is not None else None)
    n = torch.tensor([scores_local.numel()], device=scores_local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)

    def gather(t):
        pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
        pad[:t.numel()] = t.reshape(-1)
        out = torch.empty(world * mx, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad, group=group) if t.is_cuda else dist.all_gather(
            list(out.reshape(world, mx).unbind(0)), pad, group=group)
        return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)])

    return gather(scores_local), (gather(best_idx_local) if best_idx_local is not None else None)


class AsyncScoreGather:
    """The per-batch all-gather of scores / selected indices taken OFF the compute stream (the only
    collective of the path).  ``submit`` snapshots the two per-chain tensors into a send buffer on a side stream (the
    compute stream only waits for that device-to-device copy, a few microseconds, before the next batch may overwrite
    them) and issues ONE ``all_gather_into_tensor`` there; the collective then runs under the next batch's kernels.
    ``result()`` makes the caller's stream wait for the latest gather and returns (scores_all, best_idx_all) in rank
    order.  ``last_us()`` is the device time of the latest collective on this rank (CUDA events on the side stream).
    Equal shard sizes (weak scaling); on CPU tensors / without a process group it degrades to ``gather_scores``."""

    def __init__(self, n_local, device, group=None):
        self.group, self.n = group, int(n_local)
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and \
            torch.device(device).type == "cuda"
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self._fallback = None
        if not self.on:
            return
        self.stream = torch.cuda.Stream(device=device)
        self.send = torch.empty(2 * self.n, dtype=torch.float32, device=device)
        self.recv = torch.empty(self.world * 2 * self.n, dtype=torch.float32, device=device)
        self.copied, self.done = torch.cuda.Event(), torch.cuda.Event()
        self.t0, self.t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self._timed = False

    def submit(self, scores_local, best_idx_local):
        if not self.on:
            self._fallback = gather_scores(scores_local, best_idx_local, self.group, equal_sizes=True)
            return
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)  # the batch that produced the scores
        with torch.cuda.stream(self.stream):
            self.send[:self.n].copy_(scores_local.reshape(-1), non_blocking=True)
            self.send[self.n:].view(torch.int32).copy_(best_idx_local.reshape(-1), non_blocking=True)
            self.copied.record(self.stream)
            self.t0.record(self.stream)
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
            self.t1.record(self.stream)
            self.done.record(self.stream)
            self._timed = True
        cur.wait_event(self.copied)  # the next batch may overwrite scores / best_idx once they are snapshotted

    def result(self):
        if not self.on:
            return self._fallback
        torch.cuda.current_stream().wait_event(self.done)
        out = self.recv.reshape(self.world, 2, self.n)
        return out[:, 0].reshape(-1), out[:, 1].contiguous().view(torch.int32).reshape(-1)

    def last_us(self):
        if not (self.on and self._timed):
            return None
        self.t1.synchronize()
        return 1e3 * self.t0.elapsed_time(self.t1)


def reduce_metrics(partials, group=None):
    """sum a small dict of scalar partial sums over ranks (acc numerators/denominators etc.)"""
    keys = sorted(partials)
    dev = partials[keys[0]].device if isinstance(partials[keys[0]], torch.Tensor) else "cpu"
    t = torch.stack([torch.as_tensor(partials[k], dtype=torch.float64, device=dev).reshape(()) for k in keys])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if t.is_cuda:
            t = t.float()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: float(v) for k, v in zip(keys, t.tolist())}
