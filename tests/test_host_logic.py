"""CPU tier: host-side logic of the drop-in (formula compiler, parser, containers, sharding) and the
C-ABI library's load/exports.  No compute calls without a GPU."""
import ctypes as C
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import pstl_b200  # noqa: F401
from pstl_b200 import native, sharding, stl_d_lib as S, synthetic
from pstl_b200 import nusc_train as NT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = native.lib()
    hdr = open(os.path.join(ROOT, "include", "pstl.h")).read()
    declared = set(re.findall(r"\b(pstl_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for sym in sorted(declared):
        assert hasattr(L, sym), "libpstl_b200.so does not export %s" % sym
    assert set(native.EXPORTS) <= declared
    assert L.pstl_version() >= 100


def test_ops_fail_loudly_without_cuda():
    x = {"a": torch.zeros(2, 5)}
    with pytest.raises(native.PstlNativeError):
        S.Always(0, 3, S.AP(lambda q: q["a"]))(x, 100.0)
    with pytest.raises(native.PstlNativeError):
        NT.generate_trajs(torch.zeros(2, 4), torch.zeros(2, 5, 2), 0.5)


def test_compile_formula_postfix():
    a, b = S.AP(lambda x: x["a"], "a"), S.AP(lambda x: x["b"], "b")
    ops, leaves = S.compile_formula(S.Eventually(0, 4, S.Always(0, 8, S.And(a, b))))
    assert leaves == [a, b]
    assert ops == [(native.OP_SIGNAL, 0, 0), (native.OP_SIGNAL, 1, 0), (native.OP_SMIN2, 0, 0),
                   (native.OP_WIN_SMIN, 0, 8), (native.OP_WIN_SMAX, 0, 4)]
    ops, _ = S.compile_formula(S.Imply(a, b))
    assert [o[0] for o in ops] == [native.OP_SIGNAL, native.OP_NEG, native.OP_SIGNAL, native.OP_SMAX2]
    ops, leaves = S.compile_formula(S.Until(2, 5, a, b))
    assert [o[0] for o in ops] == [native.OP_SIGNAL, native.OP_WIN_SMAX, native.OP_SIGNAL, native.OP_SIGNAL,
                                   native.OP_PREFIX_SMIN, native.OP_SMIN2, native.OP_SUFFIX_SMAX, native.OP_WIN_SMIN,
                                   native.OP_SMIN2]
    assert len(leaves) == 2  # shared leaves are pushed by id, not duplicated


def test_driving_spec_compiles_to_typed_programs():
    args = NT.default_args()
    stls = NT.build_stl_cache(args)
    sizes = []
    for f in stls:
        ops, leaves = S.compile_formula(f, fused=True)
        assert not leaves and all(o[0] != native.OP_SIGNAL for o in ops)
        sizes.append(len(ops))
    assert sizes == [13, 15, 15]  # SURVEY §8(a) a12 node counts
    with pytest.raises(ValueError):
        S.compile_formula(S.Always(0, 3, S.AP(lambda x: x)), fused=True)


def test_str_matches_reference_format():
    a, b = S.AP(None, "a"), S.AP(None, "b")
    f = S.Eventually(0, 4, S.Always(0, 8, S.And(a, b)))
    assert str(f) == "♢[0:5] (◻[0:9] ((a) & (b)))"
    f.update_format("word")
    assert str(f) == "EVENTUALLY[0:5] (ALWAYS[0:9] ((a) AND (b)))"
    n0 = S.AP.n_aps
    assert str(S.AP(None)) == "AP%d" % n0 and S.AP.n_aps == n0 + 1


@pytest.mark.ref
def test_parser_matches_reference_defaults_and_overrides():
    import ref_shim
    for flags in (ref_shim.OURS_FLAGS, ref_shim.GUIDE_FLAGS):
        _, rargs = ref_shim.load(flags)
        ours = NT.generate_parser(list(flags))
        r, o = vars(rargs), vars(ours)
        for k, v in r.items():
            assert k in o, "flag %s missing" % k
            assert o[k] == v, "flag %s: %r != %r" % (k, o[k], v)
        assert set(o) - set(r) <= {"synthetic", "precision"}


def test_schedule_matches_oracle():
    from oracle import pstl_oracle as O
    args = NT.default_args()
    beta, alpha, ah = [t.cpu() for t in NT.get_diffusion_coeffs(args)]
    b2, a2, h2 = O.ddpm_schedule(100)
    assert torch.equal(beta, b2) and torch.equal(alpha, a2) and torch.equal(ah, h2)
    assert abs(beta[10].item() - 1.1061e-3) < 1e-6 and abs(ah[99].item() - 0.1946) < 1e-4


def test_iterate_list_semantics():
    kept = torch.arange(5 * 2 * 3 * 2).reshape(5, 2, 3, 2).float()
    il = NT.IterateList(100, kept)
    assert len(il) == 100
    assert torch.equal(il[-1], kept[-1]) and torch.equal(il[99], kept[4]) and torch.equal(il[95], kept[0])
    assert [t.sum().item() for t in il[-5:]] == [kept[i].sum().item() for i in range(5)]
    with pytest.raises(IndexError):
        il[50]


def test_lazy_batch_materialises_on_read_only():
    calls = []
    lb = NT.LazyBatch({"a": 1})
    lb.set_lazy("big", lambda: calls.append(1) or 42)
    assert "big" in lb and not calls
    assert lb["big"] == 42 and lb["big"] == 42 and calls == [1]


def test_augment_batch_index_conventions():
    """flat chain index n=(scene*S+sample)*3+mode on CPU tensors"""
    args = NT.default_args(n_randoms=4, sampling_size=4)
    b = synthetic.make_scene_batch(3, n_randoms=4, seed=1)
    b["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
    nb = NT.augment_batch_data(NT.LazyBatch(b), None, args, n_randoms=4)
    pack = nb["_pstl_pack"]
    assert pack.N == 36 and pack.rows_per_scene == 12
    assert pack.mode.tolist() == [0.0, 1.0, 2.0] * 12
    assert torch.equal(pack.state0[12:24], b["ego_traj"][1, 0, :4].expand(12, 4))
    assert torch.equal(pack.stlp.reshape(3, 4, 3, 6)[:, 2], b["pre_stlp"].reshape(3, 4, 3, 6)[:, 0])
    v = torch.cat([b["curr_id"], b["left_id"], b["right_id"]], -1)
    assert torch.equal(pack.valid.reshape(3, 4, 3)[:, 1], v)
    assert torch.equal(nb["neighbors_dense"], NT.dup(b["neighbor_trajs_aug"], 12))


def test_shard_bounds_cover_and_are_disjoint():
    for n, w in ((1024, 8), (10, 4), (3, 8), (65536, 8)):
        spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = synthetic.make_scene_batch(5, n_randoms=4, seed=9)
    mine = sharding.shard_batch(b, rank, world)
    lo, hi = sharding.shard_bounds(5, rank, world)
    ok = torch.equal(mine["ego_traj"], b["ego_traj"][lo:hi])
    scores = torch.arange(lo * 12, hi * 12).float()  # 12 chains per scene
    idx = torch.arange(lo * 12, hi * 12).int()
    all_s, all_i = sharding.gather_scores(scores, idx)
    ok = ok and torch.equal(all_s, torch.arange(60).float()) and torch.equal(all_i, torch.arange(60).int())
    valid = (torch.arange(lo * 12, hi * 12) % 3 != 1).float()
    n_tot, mean_v = sharding.guidance_normaliser(valid)
    ok = ok and n_tot == 60 and abs(mean_v - 2.0 / 3.0) < 1e-6
    # the side-stream gather degrades to the blocking collective on CPU tensors (equal shard sizes)
    g = sharding.AsyncScoreGather(12, "cpu")
    g.submit(torch.arange(12).float() + 100 * rank, torch.arange(12).int() + 7 * rank)
    ga, gi = g.result()
    ok = ok and torch.equal(ga, torch.cat([torch.arange(12).float() + 100 * r for r in range(world)]))
    ok = ok and torch.equal(gi, torch.cat([torch.arange(12).int() + 7 * r for r in range(world)])) and g.last_us() is None
    red = sharding.reduce_metrics({"num": torch.tensor(float(rank + 1)), "den": torch.tensor(2.0)})
    ok = ok and red == {"den": 2.0 * world, "num": world * (world + 1) / 2}
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_closed_loop_pick_masks_lane_change_modes():
    """reference nusc_sim.py:677-683"""
    g = torch.Generator().manual_seed(0)
    sc = torch.randn(64 * 3, generator=g)
    sc[5 * 3 + 1] = 50.0          # a lane-change row with the best raw score must not win
    sc[7 * 3] = sc[11 * 3] = 9.0  # tie between two mode-0 rows: first maximum
    u, tr = torch.randn(192, 20, 2, generator=g), torch.randn(192, 21, 4, generator=g)
    idx, best, cu, ct = NT.closed_loop_pick(sc, u, tr)
    cube = sc.reshape(64, 3).clone()
    cube[:, 1:3] = -10000
    assert int(idx) == int(torch.argmax(cube)) == 21 and float(best) == 9.0
    assert cu.shape == (1, 20, 2) and ct.shape == (1, 21, 4) and torch.equal(cu[0], u[21])
    assert sc[5 * 3 + 1] == 50.0  # input untouched


def test_trajopt_params_files(tmp_path):
    """save_trajopt_params / load_trajopt_params: the reference's per-sample .npy names and shapes
    (nusc_train.py:775-797, nusc_dataset.py:203-225), nothing written under --test"""
    from pstl_b200 import nusc_train as NT
    args = NT.default_args(n_randoms=4)
    args.test = False
    bs, nt = 3, args.nt
    g = torch.Generator().manual_seed(1)
    params = torch.randn(bs, 4, 3, nt, 2, generator=g)
    init = torch.randn(bs, 4, 3, nt, 2, generator=g)
    scores = torch.randn(bs, 4, 3, generator=g)
    stlp = torch.randn(bs * 4 * 3, 1, 6, generator=g)
    traj_i, ti = torch.tensor([7, 7, 123]), torch.tensor([0, 15, 2])
    d = str(tmp_path)
    NT.save_trajopt_params(init, "init", traj_i, ti, args, model_dir=d)
    NT.save_trajopt_params(params, 50, traj_i, ti, args, model_dir=d)
    NT.save_trajopt_params(scores, "scores", traj_i, ti, args, model_dir=d)
    names = NT.save_trajopt_params(params, "final", traj_i, ti, args, save_stlp=stlp, model_dir=d)
    assert names[:2] == ["params_00007_0000.npy", "params_00007_0000_stlp.npy"]
    assert "params_00123_0002_iter00050.npy" in os.listdir(d) and "scores_00007_0015.npy" in os.listdir(d)
    back = NT.load_trajopt_params(d, traj_i, ti)
    assert torch.equal(back["params"], params) and torch.equal(back["params_init"], init)
    assert torch.equal(back["tj_scores_prior"], scores) and back["pre_stlp"].shape == (bs, 4, 3, 1, 6)
    assert torch.equal(back["pre_stlp"].reshape(-1, 1, 6), stlp)
    args.test = True
    assert NT.save_trajopt_params(params, "final", traj_i, ti, args, model_dir=str(tmp_path / "none")) == []
    assert not os.path.exists(tmp_path / "none")


def test_offline_cache_round_trip(tmp_path):
    """nusc_dataset: cache.npz in the reference's layout (data[traj_i][ti][key], meta_list) written from scene batches and
    read back through the offline dataset + DataLoader, with the traj-opt files attached (nusc_train.py:153-208,
    nusc_dataset.py:109-118, 203-240)"""
    from pstl_b200 import nusc_train as NT, nusc_dataset as ND, synthetic
    args = NT.default_args(n_randoms=4, batch_size=2, num_workers=0)
    args.test = False
    b = synthetic.make_scene_batch(3, n_randoms=4, seed=9)
    b["traj_i"], b["ti"] = torch.tensor([5, 5, 9]), torch.tensor([1, 2, 1])
    saved = ND.save_cache_data({k: v for k, v in b.items() if k not in ("pre_stlp", "tj_scores_prior", "params_init")}, {})
    assert sorted(saved) == [5, 9] and sorted(saved[5]) == [1, 2] and "params" not in saved[5][1]
    path = str(tmp_path / "cache.npz")
    ND.write_cache(path, saved, [(5, ["t0", "t1", "t2"]), (9, ["u0", "u1"])])
    data, meta = ND.read_cache(path)
    assert np.array_equal(data[9][1]["ego_traj"], b["ego_traj"][2].numpy()) and meta[1][0] == 9
    pdir = str(tmp_path / "models")
    NT.save_trajopt_params(b["params_init"], "init", b["traj_i"], b["ti"], args, model_dir=pdir)
    NT.save_trajopt_params(b["tj_scores_prior"], "scores", b["traj_i"], b["ti"], args, model_dir=pdir)
    NT.save_trajopt_params(b["params"], "final", b["traj_i"], b["ti"], args, save_stlp=b["pre_stlp"].reshape(-1, 1, 6), model_dir=pdir)
    split = tmp_path / "split.txt"
    split.write_text("5 1 tok_a\n9 1 tok_b\n5 2 tok_c\n")
    loader = ND.get_dataloader(args, path, str(split), pdir, shuffle=False)
    batches = list(loader)
    assert [x["traj_i"].tolist() for x in batches] == [[5, 9], [5]]
    first = batches[0]
    for k in ("ego_traj", "neighbors_traj", "currlane_wpts", "curr_id", "gt_high_level"):
        assert torch.equal(first[k], b[k][[0, 2]].float()), k
    assert torch.equal(first["params"], b["params"][[0, 2]]) and torch.equal(first["pre_stlp"], b["pre_stlp"][[0, 2]])
    assert torch.equal(first["tj_scores_prior"], b["tj_scores_prior"][[0, 2]])
    # no files: fresh initial controls in upstream's ranges
    fresh = ND.CacheDataset(data, args)[0]
    assert fresh["params"].shape == (4, 3, args.nt, 2) and float(fresh["params"][..., 0].abs().max()) <= 0.1 * args.mul_w_max


@pytest.mark.ref
def test_offline_cache_loads_in_reference_dataset(tmp_path, monkeypatch):
    """a cache + traj-opt files written by this package load in the UNMODIFIED reference's MyDataset (--offline), sample
    for sample equal to what nusc_dataset.CacheDataset returns; and the reference's save_cache_data output reads here"""
    import ref_shim
    from pstl_b200 import nusc_dataset as ND, synthetic
    T, rargs = ref_shim.load(["-e", "e5_ddpm", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--n_randoms", "4",
                              "--params_load_path", "e1"])
    import nusc_dataset as RD
    args = NT.default_args(n_randoms=4, batch_size=2, num_workers=0)
    args.test = False
    b = synthetic.make_scene_batch(3, n_randoms=4, seed=9)
    b["traj_i"], b["ti"] = torch.tensor([5, 5, 9]), torch.tensor([1, 2, 1])
    scene = {k: v for k, v in b.items() if k not in ("pre_stlp", "tj_scores_prior", "params_init")}
    ours = ND.save_cache_data(scene, {})
    theirs = T.save_cache_data(scene, {})
    for t in theirs:
        for k in theirs[t]:
            assert sorted(theirs[t][k]) == sorted(ours[t][k])
            for key in theirs[t][k]:
                assert np.array_equal(theirs[t][k][key], ours[t][k][key]), key
    # reference layout on disk: <root>/e5/ (exp_dir_full), <root>/e1/models/ (params_load_path)
    root = tmp_path
    (root / "e5").mkdir()
    pdir = str(root / "e1" / "models")
    NT.save_trajopt_params(b["params_init"], "init", b["traj_i"], b["ti"], args, model_dir=pdir)
    NT.save_trajopt_params(b["tj_scores_prior"], "scores", b["traj_i"], b["ti"], args, model_dir=pdir)
    NT.save_trajopt_params(b["params"], "final", b["traj_i"], b["ti"], args, save_stlp=b["pre_stlp"].reshape(-1, 1, 6), model_dir=pdir)
    meta = [(5, ["t0", "t1", "t2"]), (9, ["u0", "u1"])]
    rargs.offline, rargs.exp_dir_full, rargs.generate_split_on_the_fly, rargs.test = True, str(root / "e5"), True, False
    rargs.train_ratio = 1.0
    ds_ref = RD.MyDataset(None, None, meta, ours, "train", rargs)
    ds_ref.indices = [(5, 1, "t1"), (5, 2, "t2"), (9, 1, "u1")]
    ds = ND.CacheDataset(ours, args, ds_ref.indices, pdir)
    for i in range(3):
        r, o = ds_ref[i], ds[i]
        assert set(r) == set(o), set(r) ^ set(o)
        for k in r:
            if isinstance(r[k], torch.Tensor):
                assert r[k].dtype == o[k].dtype and torch.equal(r[k], o[k]), k
            else:
                assert r[k] == o[k], k


def test_header_is_plain_c(tmp_path):
    """include/pstl.h is the drop-in boundary: it must compile as C99 on its own (no torch / C++ types)"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "pstl.h"\nint main(void) { pstl_loss_cfg c; pstl_spec_params s; (void)c; (void)s; return PSTL_LOSS_N; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_program_links_the_library(tmp_path):
    """a plain C consumer: compile against include/pstl.h, link libpstl_b200.so, call the entry points that need no GPU"""
    import shutil
    import subprocess
    from pstl_b200 import native
    if shutil.which("gcc") is None or not os.path.exists(native.LIB_PATH):
        pytest.skip("no gcc or library not built")
    src = tmp_path / "use.c"
    src.write_text('#include <stdio.h>\n#include "pstl.h"\n'
                   'int main(void) {\n'
                   '  pstl_loss_cfg c = {0};\n'
                   '  printf("%d|%zu|%d\\n", pstl_version(), pstl_refine_losses_workspace_bytes(&c),\n'
                   '         pstl_refine_losses(&c, 0, 0, 0, 0, 0, 0, 0, 0, 0));\n'
                   '  printf("%s\\n", pstl_last_error());\n  return 0;\n}\n')
    libdir = os.path.dirname(native.LIB_PATH)
    exe = tmp_path / "use"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir, "-lpstl_b200",
                        "-Wl,-rpath," + libdir, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    first, second = out.stdout.strip().split("\n")[:2]
    ver, ws, rc = first.split("|")
    assert int(ver) > 0 and int(ws) == 64 and int(rc) != 0 and "null argument" in second


@pytest.mark.ref
def test_evaluate_all_scores_matches_reference():
    """evaluate_all_scores (nusc_train.py:347-368): same keys, same score vectors in the same order"""
    import ref_shim
    T, rargs = ref_shim.load(ref_shim.OURS_FLAGS + ["--n_randoms", "8"])
    g = torch.Generator().manual_seed(3)
    bs, S_ = 6, 8
    scores = torch.randn(bs * S_ * 3, generator=g)
    labels = torch.tensor([[0.0], [1.0], [2.0], [3.0], [1.0], [0.0]])
    valid = (torch.rand(bs, 1, 3, generator=g) > 0.3).float().repeat(1, S_, 1).reshape(-1)
    ref = T.evaluate_all_scores(scores, labels, valid)
    ours = NT.evaluate_all_scores(scores, labels, valid, S_)
    assert set(ref) == set(ours)
    for k in ref:
        assert len(ref[k]) == len(ours[k]), k
        for a, b in zip(ref[k], ours[k]):
            assert torch.equal(a, b), k


@pytest.mark.ref
def test_diffusion_prep_matches_reference():
    """diffusion_prep (nusc_train.py:539-555) draws from torch's global generator in upstream's order: same seed, same
    noise, timesteps and noised commands"""
    import ref_shim
    T, rargs = ref_shim.load(["-e", "e5_ddpm", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--n_randoms", "4"])
    args = NT.default_args(flags=["-e", "e5_ddpm", "--diffusion", "--stl_weight", "0.0", "--load_stlp"], n_randoms=4)
    g = torch.Generator().manual_seed(8)
    controls = torch.randn(5, 4, 3, args.nt, 2, generator=g)
    coeffs = T.get_diffusion_coeffs(rargs)
    torch.manual_seed(123)
    r_noise, r_t, _, r_noised = T.diffusion_prep(controls, n_randoms=4, coeffs=coeffs)
    torch.manual_seed(123)
    noise, t, _, noised = NT.diffusion_prep(controls, 4, [c.cpu() for c in NT.get_diffusion_coeffs(args)], args)
    assert torch.equal(t, r_t) and torch.equal(noise, r_noise)
    np.testing.assert_allclose(noised.numpy(), r_noised.numpy(), rtol=1e-6, atol=1e-6)


@pytest.mark.ref
def test_save_trajopt_params_matches_reference(tmp_path):
    """save_trajopt_params writes the same files with the same contents as the reference's (nusc_train.py:775-797)"""
    import ref_shim
    T, rargs = ref_shim.load(["-e", "e1", "--diffusion", "--load_stlp", "--trajopt_only", "--n_randoms", "4"])
    args = NT.default_args(n_randoms=4)
    args.test = rargs.test = False
    g = torch.Generator().manual_seed(2)
    params = torch.randn(2, 4, 3, args.nt, 2, generator=g)
    stlp = torch.randn(2 * 4 * 3, 1, 6, generator=g)
    traj_i, ti = torch.tensor([3, 41]), torch.tensor([7, 0])
    d_ref, d_our = tmp_path / "ref", tmp_path / "our"
    d_ref.mkdir()
    rargs.model_dir = str(d_ref)
    for it, st in (("init", None), (12, None), ("scores", None), ("final", stlp)):
        payload = params[..., 0, 0] if it == "scores" else params
        T.save_trajopt_params(payload, it, traj_i, ti, rargs, save_stlp=st)
        NT.save_trajopt_params(payload, it, traj_i, ti, args, save_stlp=st, model_dir=str(d_our))
    assert sorted(os.listdir(d_ref)) == sorted(os.listdir(d_our)) and len(os.listdir(d_our)) == 10
    for f in os.listdir(d_ref):
        assert np.array_equal(np.load(d_ref / f), np.load(d_our / f)), f
