"""GPU tier: the time-parallel scorer of the reference's dense per-row layout (csrc/score_dense.cuh; BASELINE configs 1
and 5 through compute_stl_dense) — bit-exact against the streaming scorer it replaces for that layout, 1e-5 against the
oracle and the reference's golden scores, at the horizon / neighbour corners of config 5 and on ragged sizes."""
import os

import numpy as np
import pytest
import torch

import pstl_b200  # noqa: F401
from pstl_b200 import synthetic
from pstl_b200 import nusc_train as NT
from oracle import pstl_oracle as O
from test_gpu_parity import close, cuda
from test_gpu_flags import close_elem

pytestmark = pytest.mark.gpu


def _scores(x, idx, mask, args, kernel):
    stls = NT.build_stl_cache(args)
    os.environ["PSTL_SCORE_KERNEL"] = kernel
    try:
        _, sc, _ = NT.compute_stl_dense(x, stls, idx, mask, args)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("PSTL_SCORE_KERNEL", None)
    return sc.clone()


@pytest.mark.parametrize("n,nt,knei,seed", [(4096, 20, 8, 1008), (1537, 20, 8, 1031), (333, 50, 16, 1032), (96, 100, 32, 1022),
                                            (64, 200, 64, 1021), (515, 24, 4, 1033), (7, 20, 8, 1034)])
def test_dense_tp_equals_stream_and_oracle(n, nt, knei, seed):
    x, idx, mask = synthetic.make_dense_stl_input(n, nt=nt, n_neighbors=knei, seed=seed)
    idx = idx.clone()
    idx[5::17] = 3.0  # some outlier-mode rows (constant score 1.0, nusc_train.py:322)
    args = NT.default_args(nt=nt)
    xc = cuda(x)
    tp = _scores(xc, idx.cuda(), mask.cuda(), args, "dense")
    st = _scores(xc, idx.cuda(), mask.cuda(), args, "stream")
    assert torch.equal(tp, st), float((tp - st).abs().max())
    ref = O.stl_scores(dict(x), idx[:, 0], 100.0)
    close_elem(tp, ref, rtol=2e-5, floor=1.0, what="dense tp vs oracle")
    assert (tp[idx[:, 0].cuda() == 3] == 1.0).all()


def test_dense_tp_golden(golden_dir):
    """the reference's own compute_stl_dense scores (stl_dense.npz) through the time-parallel kernel"""
    G = np.load(os.path.join(golden_dir, "stl_dense.npz"))
    for tag, n, nt, knei, seed in (("t20k8", 192, 20, 8, 1008), ("t50k16", 48, 50, 16, 1009)):
        x, idx, mask = synthetic.make_dense_stl_input(n, nt=nt, n_neighbors=knei, seed=seed)
        tp = _scores(cuda(x), idx.cuda(), mask.cuda(), NT.default_args(nt=nt), "dense")
        close_elem(tp, G[tag + "|scores"], rtol=2e-5, floor=1.0, what=tag)


def test_dense_tp_flags_and_fallbacks(golden_dir):
    """--inline / --clip_dist / --norm_stl reach the kernel through the plan and the lane flag word; shapes it does not
    take (K*T not a multiple of 4: the 16-byte bulk copies) fall back to the streaming scorer with the same result"""
    G = np.load(os.path.join(golden_dir, "flags.npz"))
    x, idx, mask = synthetic.make_dense_stl_input(96, nt=20, n_neighbors=8, seed=1010, endcaps=True, overlap=True)
    for tag, over in (("inline", dict(inline=True)), ("inline_clip", dict(inline=True, clip_dist=True)), ("norm", dict(norm_stl=True))):
        tp = _scores(cuda(x), idx.cuda(), mask.cuda(), NT.default_args(nt=20, **over), "dense")
        close_elem(tp, G[tag + "|scores"], rtol=2e-5, floor=1.0, what=tag)
    x, idx, mask = synthetic.make_dense_stl_input(130, nt=31, n_neighbors=3, seed=1023)
    a = _scores(cuda(x), idx.cuda(), mask.cuda(), NT.default_args(nt=31), "dense")
    b = _scores(cuda(x), idx.cuda(), mask.cuda(), NT.default_args(nt=31), "stream")
    assert torch.equal(a, b)


@pytest.mark.parametrize("rows_per_block", [1, 3, 12])
def test_dense_tp_block_shapes(rows_per_block):
    """the rows-per-block / chunk parameters change the schedule, never the result"""
    x, idx, mask = synthetic.make_dense_stl_input(1000, nt=20, n_neighbors=8, seed=1035)
    args = NT.default_args(nt=20)
    xc = cuda(x)
    ref = _scores(xc, idx.cuda(), mask.cuda(), args, "stream")
    os.environ["PSTL_DENSE_ROWS"] = str(rows_per_block)
    os.environ["PSTL_DENSE_CHUNK_KB"] = "4"  # forces the chunked, double-buffered path (2 neighbours per chunk at R=3)
    try:
        tp = _scores(xc, idx.cuda(), mask.cuda(), args, "dense")
    finally:
        os.environ.pop("PSTL_DENSE_ROWS", None)
        os.environ.pop("PSTL_DENSE_CHUNK_KB", None)
    assert torch.equal(tp, ref)


def test_dense_tp_full_size_properties():
    """config 5's first cell at its full size (1,000,128 dense rows): the scores are those of the streaming scorer on a
    random sample of rows, rows of mode 3 score 1.0, every score is finite"""
    from bench import _dense_rows_on_device
    n = 1000128
    x, idx, mask = _dense_rows_on_device(n, 20, 8, torch.device("cuda"), 77)
    args = NT.default_args(nt=20)
    tp = _scores(x, idx, mask, args, "dense")
    assert torch.isfinite(tp).all()
    g = torch.Generator().manual_seed(3)
    sel = torch.randperm(n, generator=g)[:20000].cuda()
    sub = {k: v[sel].contiguous() for k, v in x.items()}
    st = _scores(sub, idx[sel].contiguous(), mask[sel].contiguous(), args, "stream")
    assert torch.equal(tp[sel], st)
