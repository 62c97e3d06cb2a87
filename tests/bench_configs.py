"""Auxiliary measurements of the other BASELINE configs (not the driver's bench line).
   python tests/bench_configs.py guidance [scenes]     # config 3 shape: K=10, guidance last 10 steps, n_rolls 3
   python tests/bench_configs.py dense [n] [T] [Knei]  # config 1/5 shape: compute_stl_dense on dense rows
   python tests/bench_configs.py trajopt [scenes] [iters]  # trajectory-optimisation iterations
   python tests/bench_configs.py losses [scenes]       # RefineNet training losses, value + gradient (§8(f) item 4)
   python tests/bench_configs.py train [scenes] [steps]  # full --rect_head training iterations (sampler .. Adam step)
   python tests/bench_configs.py train_ddpm [scenes] [steps]  # denoiser training iterations (README step 1)
   python tests/bench_configs.py sampler [scenes] [engine]    # the bf16 sampler alone (99 reverse steps), tcgen05 engine 1 | 2
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pstl_b200  # noqa: E402,F401
from pstl_b200 import synthetic  # noqa: E402
from pstl_b200 import nusc_train as NT  # noqa: E402
from pstl_b200.nusc_model import Net  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def guidance(scenes):
    args = NT.default_args(NT.GUIDANCE_FLAGS, precision="bf16")
    net = Net(args)
    net.load_state_dict(synthetic.make_weights(1007))
    net = net.cuda()
    b = {k: v.cuda() for k, v in synthetic.make_scene_batch(scenes, seed=3).items()}
    stls, co = NT.build_stl_cache(args), NT.get_diffusion_coeffs(args)
    ms = timeit(lambda: NT.sample_and_score(net, b, stls, co, args), reps=3, warm=1)
    n = scenes * 64 * 3
    out = NT.sample_and_score(net, b, stls, co, args)
    print("config3 (Ours+guidance, K=10, n_rolls=3): scenes=%d chains=%d  %.2f ms/batch  %.3g chains/s  acc=%.3f"
          % (scenes, n, ms, n / ms * 1e3, out["acc"].item()))


def sampler(scenes, engine):
    # engine 1 | 2: the bf16 engines; 3: the split-operand (f16x3, fp32-grade) engine; 4: the one-SM engine on fp16 operands; 0: fp32 SIMT
    args = NT.default_args(precision={0: "fp32", 3: "f16x3", 4: "f16"}.get(engine, "bf16"), tc_engine=engine if engine in (1, 2) else 0)
    net = Net(args)
    net.load_state_dict(synthetic.make_weights(1007))
    net = net.cuda()
    b = NT.LazyBatch({k: v.cuda() for k, v in synthetic.make_scene_batch(scenes, seed=3).items()})
    b["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
    S_ = args.n_randoms
    b = NT.augment_batch_data(b, None, args, n_randoms=S_)
    n = scenes * S_ * 3
    noise = torch.empty((n, 40), device="cuda")
    co = NT.get_diffusion_coeffs(args)
    with torch.no_grad():
        feat = net.encode_feat(b)
    ms = timeit(lambda: NT.diffusion_rollout(noise, net, b, b["highlevel_dense"], feat, args, co, n_randoms=S_), reps=5, warm=2)
    fl = 99 * 2 * (47 * 256 + 256 * 256 + 256 * 40) * n
    print("sampler (engine %d): scenes=%d chains=%d  %.3f ms  %.1f TFLOP/s (minimal count)" % (engine, scenes, n, ms, fl / ms / 1e9))


def dense(n, T, K):
    args = NT.default_args(nt=T)
    x, idx, mask = synthetic.make_dense_stl_input(n, nt=T, n_neighbors=K, seed=7)
    xc = {k: v.cuda() for k, v in x.items()}
    idx, mask = idx.cuda(), mask.cuda()
    stls = NT.build_stl_cache(args)
    ms = timeit(lambda: NT.compute_stl_dense(xc, stls, idx, mask, args))
    by = 16 * T + 28 * K * T + 36 * 15 + 24 + 8 + 4
    print("dense compute_stl_dense: n=%d T=%d Knei=%d  %.3f ms  %.3g traj/s  %.1f GB/s algorithmic (%d B/traj)"
          % (n, T, K, ms, n / ms * 1e3, n * by / ms / 1e6, by))


def sweep(T, K, scenes=512):
    """config 5 shape: horizon T, K neighbours; scene-indexed pack (scenes x 192 rows) and dense per-row layout"""
    args = NT.default_args(nt=T)
    S = 64
    b = {k: v.cuda() for k, v in synthetic.make_scene_batch(scenes, nt=T, n_neighbors=K, n_randoms=S, seed=11).items()}
    nb = NT.LazyBatch({k: b[k] for k in ("ego_traj", "neighbors", "currlane_wpts", "leftlane_wpts", "rightlane_wpts",
                                         "curr_id", "left_id", "right_id", "gt_high_level", "pre_stlp")})
    nb["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
    pack = NT.augment_batch_data(nb, None, args, n_randoms=S)["_pstl_pack"]
    progs = NT._fused_programs(NT.build_stl_cache(args), T)
    g = torch.Generator().manual_seed(1)
    u = ((torch.rand(pack.N, T, 2, generator=g) * 2 - 1) * torch.tensor([0.05, 2.0])).cuda()
    ms = timeit(lambda: NT.score_pack(pack, u, args, progs), reps=3, warm=1)
    print("sweep T=%d K=%d scene-indexed: %d trajectories  %.3f ms  %.3g traj/s" % (T, K, pack.N, ms, pack.N / ms * 1e3))
    n = min(pack.N, int(6e9 // (16 * T + 28 * K * T + 600)))
    x, idx, mask = synthetic.make_dense_stl_input(n, nt=T, n_neighbors=K, seed=7)
    xc = {k: v.cuda() for k, v in x.items()}
    stls = NT.build_stl_cache(args)
    ms = timeit(lambda: NT.compute_stl_dense(xc, stls, idx.cuda(), mask.cuda(), args), reps=3, warm=1)
    by = 16 * T + 28 * K * T + 36 * 15 + 24 + 8 + 4
    print("sweep T=%d K=%d dense rows   : %d trajectories  %.3f ms  %.3g traj/s  %.1f GB/s algorithmic (%d B/traj)"
          % (T, K, n, ms, n / ms * 1e3, n * by / ms / 1e6, by))


def trajopt(scenes, iters):
    """trajectory optimisation (nusc_train.py:1303-1325): iterations/s on scenes x 64 x 3 stored control sequences"""
    args = NT.default_args()
    b = {k: v.cuda() for k, v in synthetic.make_scene_batch(scenes, seed=5).items()}
    nb = NT.LazyBatch(dict(b))
    nb["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
    nb = NT.augment_batch_data(nb, None, args)
    stls = NT.build_stl_cache(args)
    ms = timeit(lambda: NT.trajopt(nb, stls, args, iters=iters), reps=2, warm=1)
    n = scenes * 64 * 3
    print("trajopt: scenes=%d trajectories=%d  %d Adam iterations in %.1f ms = %.3f ms/iteration  (%.3g trajectory-gradients/s)"
          % (scenes, n, iters, ms, ms / iters, n * iters / ms * 1e3))


def losses(scenes):
    """compute_policy_loss of the --rect_head --diverse_loss step (nusc_train.py:370-478): rollout -> scores -> loss
    terms and the gradient w.r.t. rect_controls, against the CPU oracle's autograd on a slice"""
    from oracle import pstl_oracle as O
    args = NT.default_args(stl_weight=0.5, rect_reg_loss=0.1)
    S = args.n_randoms
    b = {k: v.cuda() for k, v in synthetic.make_scene_batch(scenes, seed=5).items()}
    nb = NT.LazyBatch(dict(b))
    nb["neighbor_trajs_aug"] = b["neighbors_traj"][..., :7]
    nb = NT.augment_batch_data(nb, b["pre_stlp"].reshape(scenes, S, 3, 6)[:, 0, 0], args)
    stls = NT.build_stl_cache(args)
    N = scenes * S * 3
    states = b["ego_traj"][:, 0, :4].unsqueeze(1).repeat(1, S * 3, 1).reshape(N, 4)
    nn = nb["params"].reshape(N, args.nt, 2).contiguous()
    rect = (nn + 0.02 * torch.randn_like(nn)).requires_grad_()
    zeros = torch.zeros(N, args.nt * 2, device="cuda")
    nn_trajs = NT.generate_trajs(states, nn, args.dt)

    def step():
        rect.grad = None
        rt = NT.generate_trajs(states, rect, args.dt)
        ex = (None, zeros, nb["highlevel_dense"], zeros[:, 0], nb["valids_dense"].reshape(-1), 0, zeros, nn, None, rect)
        rd, _ = NT.compute_policy_loss(nb, None, stls, nn_trajs, rt, None, args, diffusion_extras=ex)
        rd["loss"].backward()
        return rd

    ms = timeit(step, reps=5, warm=2)
    cfg = NT.loss_cfg(args, scenes, S)
    sc, vl = step()["scores"].detach(), nb["valids_dense"].reshape(-1).float().contiguous()
    r2, n2 = rect.detach().reshape(N, -1).contiguous(), nn.reshape(N, -1).contiguous()
    out, dr, ds = torch.empty(8, device="cuda"), torch.empty_like(r2), torch.empty_like(sc)
    L, C = NT._nv.lib(), NT._nv.C
    ws = torch.empty(L.pstl_refine_losses_workspace_bytes(C.byref(cfg)), dtype=torch.uint8, device="cuda")
    fp, st = NT._nv.fptr, NT._nv.stream()
    k_ms = timeit(lambda: L.pstl_refine_losses(C.byref(cfg), fp(r2), fp(n2), fp(sc), fp(vl), fp(out), fp(dr), fp(ds),
                                               NT._nv.ptr(ws), st), reps=20, warm=3)
    sub = min(scenes, 64)
    n_sub = sub * S * 3
    rc, scc = r2[:n_sub].cpu().reshape(n_sub, args.nt, 2).requires_grad_(), sc[:n_sub].cpu().requires_grad_()
    t0 = time.time()
    o = O.refine_losses(rc, n2[:n_sub].cpu(), scc, vl[:n_sub].cpu(), n_scenes=sub, S=S, nt=args.nt, n_shards=args.n_shards,
                        stl_weight=0.5, rect_reg_loss=0.1, stl_nn_thres=args.stl_nn_thres)
    torch.autograd.grad(o["loss"], [rc, scc])
    cpu_ms = (time.time() - t0) * 1e3 * scenes / sub
    print("losses: scenes=%d rows=%d groups=%d  training-step loss fwd+bwd (rollout, scorer, loss kernel) %.3f ms; "
          "pstl_refine_losses alone %.3f ms (%.3g groups/s); oracle autograd (loss terms only, %d threads, scaled from "
          "%d scenes) %.1f ms" % (scenes, N, scenes * 3 * args.n_shards, ms, k_ms, scenes * 3 * args.n_shards / k_ms * 1e3,
                                  torch.get_num_threads(), sub, cpu_ms))


def train_ddpm(scenes, steps):
    """README step 1: denoiser training iterations (diffusion_prep -> eps with a timestep per row -> loss -> backward
    into policy_net and the encoders -> Adam over all parameters)"""
    args = NT.default_args(flags=["-e", "e5_ddpm", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--skip_nusc_load"])
    net = Net(args)
    net.load_state_dict({k: v for k, v in synthetic.make_weights(1007).items() if not k.startswith(("rect_net", "merge_net"))})
    net = net.cuda()
    b = {k: v.cuda() for k, v in synthetic.make_scene_batch(scenes, seed=5).items()}
    coeffs = NT.get_diffusion_coeffs(args)
    opt = torch.optim.Adam(net.parameters(), lr=args.lr)
    log = []
    ms = timeit(lambda: log.append(NT.train_step_ddpm(net, b, coeffs, args, opt)["loss"].detach()), reps=steps, warm=2)
    n = scenes * args.n_randoms * 3
    print("train_ddpm: scenes=%d rows=%d  %.2f ms per iteration (%.3g rows/s); loss %s"
          % (scenes, n, ms, n / ms * 1e3, " ".join("%.4f" % float(v) for v in log[:2] + log[-2:])))
    from oracle import pstl_oracle as O
    sub = min(scenes, 16)
    bc = synthetic.make_scene_batch(sub, seed=5)
    noise, t, _, noised = NT.diffusion_prep(bc["params"], args.n_randoms, [c.cpu() for c in coeffs], args)
    t0 = time.time()
    O.ddpm_train_step(synthetic.make_weights(1007), bc, noise, t, noised, args.n_randoms, args.nt)
    cpu = time.time() - t0
    print("train_ddpm: oracle (torch CPU, %d threads) forward + backward on %d scenes: %.2f s -> %.1f s per %d-scene "
          "iteration" % (torch.get_num_threads(), sub, cpu, cpu * scenes / sub, scenes))


def train(scenes, steps):
    """README "Ours" training stage (nusc_train.py:1352-1427, 1523-1525): iterations/s of sampler -> best-of-5 ->
    RefineNet -> rollout -> scorer -> losses -> backward -> Adam over rect_net"""
    args = NT.default_args(precision="bf16", stl_weight=0.5, rect_reg_loss=0.1)
    net = Net(args)
    net.load_state_dict(synthetic.make_weights(1007))
    net = net.cuda().train()
    b = {k: v.cuda() for k, v in synthetic.make_scene_batch(scenes, seed=5).items()}
    stls = NT.build_stl_cache(args)
    coeffs = NT.get_diffusion_coeffs(args)
    opt = torch.optim.Adam(net.rect_net.parameters(), lr=args.lr)
    log = []

    def it():
        log.append(NT.train_step_rect(net, b, stls, coeffs, args, opt)["loss"].detach())

    ms = timeit(it, reps=steps, warm=2)
    n = scenes * args.n_randoms * 3
    print("train: scenes=%d chains=%d  %.2f ms per iteration (%.3g chains/s); loss %s"
          % (scenes, n, ms, n / ms * 1e3, " ".join("%.4f" % float(v) for v in log[:2] + log[-2:])))
    # the CPU oracle for the part after the sampler (RefineNet, rollout, scores, losses, backward), on a slice
    from oracle import pstl_oracle as O
    sub = min(scenes, 8)
    bc = synthetic.make_scene_batch(sub, seed=5)
    kw = dict(n_scenes=sub, S=args.n_randoms, nt=args.nt, n_shards=args.n_shards, diverse_loss=True, diverse_detach=False,
              w_max=args.mul_w_max, a_max=args.mul_a_max, stl_nn_thres=args.stl_nn_thres, stl_weight=0.5,
              diversity_scale=args.diversity_scale, diversity_weight=args.diversity_weight, rect_reg_loss=0.1,
              extra_rect_reg=0.0)
    t0 = time.time()
    O.refine_train_step(synthetic.make_weights(1007), bc, torch.zeros(sub, 224), bc["params"].reshape(-1, args.nt, 2),
                        args.dt, **kw)
    cpu = time.time() - t0
    print("train: oracle (torch CPU, %d threads) RefineNet + losses + backward on %d scenes: %.2f s -> %.1f s per %d-scene "
          "iteration, without the sampler" % (torch.get_num_threads(), sub, cpu, cpu * scenes / sub, scenes))


if __name__ == "__main__":
    if sys.argv[1] == "train_ddpm":
        train_ddpm(int(sys.argv[2]) if len(sys.argv) > 2 else 1024, int(sys.argv[3]) if len(sys.argv) > 3 else 5)
    elif sys.argv[1] == "train":
        train(int(sys.argv[2]) if len(sys.argv) > 2 else 1024, int(sys.argv[3]) if len(sys.argv) > 3 else 5)
    elif sys.argv[1] == "losses":
        losses(int(sys.argv[2]) if len(sys.argv) > 2 else 1024)
    elif sys.argv[1] == "sweep":
        sweep(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 512)
    elif sys.argv[1] == "trajopt":
        trajopt(int(sys.argv[2]) if len(sys.argv) > 2 else 256, int(sys.argv[3]) if len(sys.argv) > 3 else 100)
    elif sys.argv[1] == "sampler":
        sampler(int(sys.argv[2]) if len(sys.argv) > 2 else 1024, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    elif sys.argv[1] == "guidance":
        guidance(int(sys.argv[2]) if len(sys.argv) > 2 else 256)
    else:
        dense(*(int(a) for a in (sys.argv[2:5] + ["4096", "20", "8"][len(sys.argv) - 2:])))
