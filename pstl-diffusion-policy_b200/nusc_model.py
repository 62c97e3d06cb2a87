"""Drop-in for the reference's ``nusc_model.Net`` on the diffusion + RefineNet path
(reference nusc_model.py:8-235): same constructor, sub-module names (so reference checkpoints
load with ``load_state_dict``), ``forward`` / ``rect_forward`` signatures.

The eps-MLP, RefineNet and (without autograd) the per-scene encoders run in libpstl_b200.so; with autograd
on the encoders are the plain nn.Sequential stacks.  VAE / BC / init-hint variants are out of scope and raise.
"""
import torch
import torch.nn as nn

from . import native as _nv


def build_relu_nn(input_dim, output_dim, hiddens, activation_fn=nn.ReLU, last_fn=None):
    """Linear/ReLU stack with the reference's state_dict key layout ``*.0/2/4`` (reference utils.py:91-101)."""
    dims = [input_dim] + list(hiddens) + [output_dim]
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        layers.append(activation_fn())
    if last_fn is not None:
        layers[-1] = last_fn()
    else:
        del layers[-1]
    return nn.Sequential(*layers)


def normalize_xyth(state, base, valid=None, no_theta=False):
    """ego-frame transform (reference nusc_model.py:238-263)."""
    assert len(state.shape) == len(base.shape) and state.shape[0] == base.shape[0]
    x, y = state[..., 0], state[..., 1]
    bx, by, bth = base[..., 0], base[..., 1], base[..., 2]
    if valid is not None:
        xt, yt = x - bx * valid, y - by * valid
    else:
        xt, yt = x - bx, y - by
    xr = xt * torch.cos(bth) + yt * torch.sin(bth)
    yr = -xt * torch.sin(bth) + yt * torch.cos(bth)
    if no_theta:
        return torch.stack([xr, yr], dim=-1)
    th = state[..., 2]
    tr = th - bth * valid if valid is not None else th - bth
    return torch.stack([xr, yr, tr], dim=-1)


class _RectForward(torch.autograd.Function):
    """Net.rect_forward with gradients for rect_net's parameters: pstl_refine (fp32 handle) forward,
    pstl_refine_backward for (weight, bias) of layers 0, 2, 4."""

    @staticmethod
    def forward(ctx, net, scene_feat, hl, stlp, u0, scores, *params):
        a = net.args
        n, bs = u0.shape[0], scene_feat.shape[0]
        out = torch.empty((n, a.nt, 2), dtype=torch.float32, device=u0.device)
        handle = net.native_handle("fp32")
        L = _nv.lib()
        ws, ctx.epoch = _nv.activation_workspace(L.pstl_refine_backward_workspace_bytes(handle, n, bs), u0.device)
        ctx.ws = ws
        _nv.check(L.pstl_refine(handle, _nv.fptr(scene_feat), bs, n // bs, _nv.fptr(hl), _nv.fptr(stlp), _nv.fptr(u0),
                                _nv.fptr(scores), n, a.n_randoms, a.n_shards, _nv.C.c_float(a.mul_w_max),
                                _nv.C.c_float(a.mul_a_max), int(bool(a.clip_rect)), _nv.fptr(out), _nv.ptr(ws),
                                _nv.stream()), "pstl_refine")
        ctx.net = net
        ctx.save_for_backward(scene_feat, hl, stlp, u0, scores)
        ctx.shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, g):
        net, a = ctx.net, ctx.net.args
        scene_feat, hl, stlp, u0, scores = ctx.saved_tensors
        n, bs = u0.shape[0], scene_feat.shape[0]
        handle = net.native_handle("fp32")
        L = _nv.lib()
        grads = [torch.empty(s, dtype=torch.float32, device=u0.device) for s in ctx.shapes]
        reuse = _nv.activations_valid(ctx.epoch)  # no other training forward has written the buffer since
        ws = ctx.ws if reuse else _nv.workspace(L.pstl_refine_backward_workspace_bytes(handle, n, bs), u0.device, "denoiser")
        _nv.check(L.pstl_refine_backward(handle, _nv.fptr(scene_feat), bs, n // bs, _nv.fptr(hl), _nv.fptr(stlp),
                                         _nv.fptr(u0), _nv.fptr(scores), n, a.n_randoms, a.n_shards,
                                         _nv.C.c_float(a.mul_w_max), _nv.C.c_float(a.mul_a_max), int(bool(a.clip_rect)),
                                         _nv.fptr(_nv.f32(g.reshape(n, a.nt * 2))), *[_nv.fptr(t) for t in grads],
                                         reuse, _nv.ptr(ws), _nv.stream()), "pstl_refine_backward")
        return (None, None, None, None, None, None) + tuple(grads)


class _EpsRows(torch.autograd.Function):
    """Net.forward for training: eps with one timestep per row (pstl_denoiser_eps_rows) and its backward into
    policy_net's parameters and the scene feature (pstl_denoiser_eps_backward)."""

    @staticmethod
    def forward(ctx, net, scene_feat, hl, stlp, x, temb_rows, *params):
        n, bs = x.shape[0], scene_feat.shape[0]
        handle = net.native_handle("fp32")
        L = _nv.lib()
        eps = torch.empty_like(x)
        ws, ctx.epoch = _nv.activation_workspace(L.pstl_refine_backward_workspace_bytes(handle, n, bs), x.device)
        ctx.ws = ws
        sf = _nv.f32(scene_feat.detach())
        _nv.check(L.pstl_denoiser_eps_rows(handle, _nv.fptr(sf), bs, n // bs, _nv.fptr(hl), _nv.fptr(stlp), _nv.fptr(x), n,
                                           _nv.fptr(temb_rows), _nv.fptr(eps), _nv.ptr(ws), _nv.stream()),
                  "pstl_denoiser_eps_rows")
        ctx.net = net
        ctx.save_for_backward(sf, hl, stlp, x, temb_rows)
        ctx.shapes = [p.shape for p in params]
        return eps

    @staticmethod
    def backward(ctx, g):
        net = ctx.net
        sf, hl, stlp, x, temb_rows = ctx.saved_tensors
        n, bs = x.shape[0], sf.shape[0]
        handle = net.native_handle("fp32")
        L = _nv.lib()
        grads = [torch.empty(s, dtype=torch.float32, device=x.device) for s in ctx.shapes]
        d_feat = torch.empty_like(sf) if ctx.needs_input_grad[1] else None
        reuse = _nv.activations_valid(ctx.epoch)  # no other training forward has written the buffer since
        ws = ctx.ws if reuse else _nv.workspace(L.pstl_refine_backward_workspace_bytes(handle, n, bs), x.device, "denoiser")
        _nv.check(L.pstl_denoiser_eps_backward(handle, _nv.fptr(sf), bs, n // bs, _nv.fptr(hl), _nv.fptr(stlp), _nv.fptr(x), n,
                                               _nv.fptr(temb_rows), _nv.fptr(_nv.f32(g)), *[_nv.fptr(t) for t in grads],
                                               _nv.fptr(d_feat), reuse, _nv.ptr(ws), _nv.stream()),
                  "pstl_denoiser_eps_backward")
        return (None, d_feat, None, None, None, None) + tuple(grads)


class Net(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        if not args.diffusion or getattr(args, "vae", False) or getattr(args, "bc", False) or \
                getattr(args, "use_init_hint", False):
            raise NotImplementedError("pstl_b200.Net builds the --diffusion model only")
        self.output_dim = args.nt * 2
        self.feat_dim = feat_dim = 32
        self.stlp_dim = stlp_dim = 6
        self.lane_dim = 3
        self.n_segs = args.n_segs
        self.time_dim = 32
        self.ego_encoder = build_relu_nn(6, feat_dim, args.hiddens)
        self.neighbor_encoder = build_relu_nn(7, feat_dim, args.hiddens)
        self.lane_encoder = build_relu_nn(self.n_segs * self.lane_dim, feat_dim, args.hiddens)
        latent_dim = args.nt * 2 + self.time_dim + 1 + stlp_dim
        self.policy_net = build_relu_nn(latent_dim + feat_dim * 7, args.nt * 2, args.hiddens)
        if args.rect_head:
            extra_in_dim = 0
            if args.diverse_loss and not args.no_arch and args.diverse_fuse_type == "cat":
                extra_in_dim += args.nt * 2  # reference nusc_model.py:40-41: [init | pooled] side by side
            if args.diverse_loss:
                self.merge_net = build_relu_nn(args.nt * 2, args.nt * 2, [32, 32])
            self.rect_net = build_relu_nn(latent_dim - self.time_dim + feat_dim * 7 + extra_in_dim, args.nt * 2,
                                          args.rect_hiddens)
        self._handles = {}

    # --- native handle -------------------------------------------------------------------
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate_native(self):
        """Call after editing parameters through ``p.data`` (which does not bump the version counter autograd keeps):
        the next ``native_handle`` re-derives the library's weight copies."""
        for prec, c in list(self._handles.items()):
            self._handles[prec] = ((c[0][0], None), c[1], c[2], c[3])

    def native_handle(self, precision="fp32"):
        """The handle of ``_native_handle`` with ``args.tc_engine`` applied (0 automatic, 1 one-SM tcgen05 engine, 2 CTA-pair
        engine; ``pstl_denoiser_set_engine``)."""
        h = self._native_handle(precision)
        _nv.check(_nv.lib().pstl_denoiser_set_engine(h, int(getattr(self.args, "tc_engine", 0) or 0)), "pstl_denoiser_set_engine")
        return h

    def _native_handle(self, precision="fp32"):
        """pstl_denoiser_t over this module's parameters.  The handle points at the parameter storage; an in-place update
        (optimiser step: same storage, new version) only refreshes the library's own copies — the hoisted first-layer
        blocks and the bf16 operand images — on the current stream (``pstl_denoiser_refresh``: no allocation, no
        synchronisation); parameters that moved to other storage rebuild the handle."""
        prec = {"fp32": _nv.PRECISION_FP32, "bf16": _nv.PRECISION_BF16, "f16x3": _nv.PRECISION_F16X3, "f16": _nv.PRECISION_F16}[precision]
        ps = list(self.parameters())
        ptrs, vers = tuple(p.data_ptr() for p in ps), tuple(p._version for p in ps)
        cached = self._handles.get(prec)
        if cached is not None and cached[0] == (ptrs, vers):
            return cached[1]
        if cached is not None and cached[0][0] == ptrs and cached[3]:
            _nv.check(_nv.lib().pstl_denoiser_refresh(cached[1], _nv.stream()), "pstl_denoiser_refresh")
            self._handles[prec] = ((ptrs, vers), cached[1], cached[2], True)
            return cached[1]
        if len(self.args.hiddens) != 2 or len(getattr(self.args, "rect_hiddens", [0, 0])) != 2 or \
                self.args.hiddens[0] != self.args.hiddens[1]:
            raise NotImplementedError("native denoiser needs two equal hidden layers")
        w = _nv.Weights()
        keep = []
        in_place = [True]  # the handle reads the parameters' own storage (fp32, contiguous): refreshable

        def put(prefix, seq):
            for li in (0, 2, 4):
                for kind in ("weight", "bias"):
                    t = getattr(seq[li], kind).detach()
                    _nv.require_cuda(t, "model parameters")
                    tc = t.to(torch.float32).contiguous()
                    in_place[0] = in_place[0] and tc.data_ptr() == t.data_ptr()
                    t = tc
                    keep.append(t)
                    setattr(w, "%s%d_%s" % (prefix, li, kind[0]), t.data_ptr())

        put("p", self.policy_net)
        w.merge_hidden = 32
        if hasattr(self, "rect_net") and self._rect_native():
            put("r", self.rect_net)
            if hasattr(self, "merge_net"):
                put("m", self.merge_net)
        w.hidden = self.args.hiddens[0]
        w.rect_hidden = self.args.rect_hiddens[0] if hasattr(self, "rect_net") else self.args.hiddens[0]
        w.feat_dim, w.time_dim, w.T = self.feat_dim * 7, self.time_dim, self.args.nt
        h = _nv.C.c_void_p()
        _nv.check(_nv.lib().pstl_denoiser_create(_nv.C.byref(w), prec, _nv.C.byref(h)), "pstl_denoiser_create")
        if cached is not None:
            _nv.lib().pstl_denoiser_destroy(cached[1])
        self._handles[prec] = ((ptrs, vers), h, keep, in_place[0])
        return h

    def pos_encoding(self, t, channels):
        """sinusoidal embedding (reference nusc_model.py:48-53)."""
        inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2, device=t.device).float() / channels))
        a = torch.sin(t.repeat(1, channels // 2) * inv_freq)
        b = torch.cos(t.repeat(1, channels // 2) * inv_freq)
        return torch.cat([a, b], dim=-1)

    def time_table(self, steps, device):
        """(steps, time_dim) table of pos_encoding(t), t = 0..steps-1, built on the host exactly as the
        reference builds each row, then uploaded once."""
        key = (steps, str(device))
        cache = self.__dict__.setdefault("_temb_cache", {})
        if key not in cache:
            t = torch.arange(steps, dtype=torch.long).reshape(steps, 1)
            cache[key] = self.pos_encoding(t, self.time_dim).to(device=device, dtype=torch.float32).contiguous()
        return cache[key]

    # --- encoders (reference nusc_model.py:55-95) ----------------------------------------
    def encode_feat(self, nn_input, ext=None):
        bs = nn_input["ego_traj"].shape[0]
        if nn_input["ego_traj"].is_cuda and not torch.is_grad_enabled():
            return self._encode_feat_native(nn_input)
        ego = nn_input["ego_traj"][:, 0]
        ego_un = ego.unsqueeze(1)
        neis_ = nn_input["neighbors"]
        neis_xyth = normalize_xyth(neis_[..., 1:4], ego_un, neis_[..., 0])
        neis_input = torch.cat([neis_[..., 0:1], neis_xyth, neis_[..., 4:7]], dim=-1)
        lanes = torch.stack([normalize_xyth(nn_input["%slane_wpts" % k], ego_un, nn_input["%s_id" % k])
                             for k in ("curr", "left", "right")], dim=1)
        lanes_input = torch.cat([lanes[..., 0:1, :], lanes[..., 1:, :] - lanes[..., :-1, :]], dim=-2)
        lanes_input = lanes_input.reshape(bs, 3, lanes.shape[-2] * self.lane_dim)
        ego_input = torch.cat([normalize_xyth(ego[..., :3], ego[..., :3]), ego[..., 3:]], dim=-1)
        ego_feat = self._mlp(self.ego_encoder, ego_input)
        nei_feat = self._mlp(self.neighbor_encoder, neis_input)
        nei_feat = torch.cat([torch.min(nei_feat, dim=1)[0], torch.mean(nei_feat, dim=1), torch.max(nei_feat, dim=1)[0]],
                             dim=-1)
        lanes_feat = self._mlp(self.lane_encoder, lanes_input).reshape(bs, -1)
        return torch.cat([ego_feat, nei_feat, lanes_feat], dim=-1)

    def _encode_feat_native(self, nn_input):
        """encode_feat on the GPU without autograd: one launch builds the three encoder inputs (ego-frame transform,
        lane differences), the MLPs run on pstl_linear, one launch pools the neighbours and concatenates."""
        ego_traj = _nv.f32(nn_input["ego_traj"])
        bs, K = ego_traj.shape[0], nn_input["neighbors"].shape[1]
        nei = _nv.f32(nn_input["neighbors"])
        lanes = [_nv.f32(nn_input["%slane_wpts" % k]) for k in ("curr", "left", "right")]
        ids = [_nv.f32(nn_input["%s_id" % k].reshape(bs)) for k in ("curr", "left", "right")]
        nseg = lanes[0].shape[1]
        dev = ego_traj.device
        ego_in = torch.empty((bs, 6), dtype=torch.float32, device=dev)
        nei_in = torch.empty((bs * K, 7), dtype=torch.float32, device=dev)
        lane_in = torch.empty((bs * 3, nseg * self.lane_dim), dtype=torch.float32, device=dev)
        L = _nv.lib()
        _nv.check(L.pstl_encoder_inputs(_nv.fptr(ego_traj), ego_traj.shape[1] * ego_traj.shape[2], _nv.fptr(nei),
                                        _nv.fptr(lanes[0]), _nv.fptr(lanes[1]), _nv.fptr(lanes[2]), _nv.fptr(ids[0]),
                                        _nv.fptr(ids[1]), _nv.fptr(ids[2]), bs, K, nseg, _nv.fptr(ego_in),
                                        _nv.fptr(nei_in), _nv.fptr(lane_in), _nv.stream()), "pstl_encoder_inputs")
        ego_f, nei_f, lane_f = self._mlp_batch([(self.ego_encoder, ego_in), (self.neighbor_encoder, nei_in),
                                                (self.lane_encoder, lane_in)])
        F = ego_f.shape[-1]
        feat = torch.empty((bs, 7 * F), dtype=torch.float32, device=dev)
        _nv.check(L.pstl_encoder_pool(_nv.fptr(ego_f), _nv.fptr(nei_f), _nv.fptr(lane_f), bs, K, F, _nv.fptr(feat),
                                      _nv.stream()), "pstl_encoder_pool")
        return feat

    def _mlp_batch(self, jobs):
        """several encoder MLPs in one launch (pstl_mlp3_batch) when they have the built shape, else one by one"""
        ok = all(len(seq) == 5 and seq[0].out_features == 256 and seq[2].out_features == 256 and
                 seq[0].in_features <= 256 and seq[4].out_features <= 32 for seq, _ in jobs)
        if not ok or len(jobs) > 4:
            return [self._mlp(seq, x) for seq, x in jobs]
        probs = (_nv.Mlp3Problem * len(jobs))()
        keep, outs = [], []
        for i, (seq, x) in enumerate(jobs):
            h = _nv.f32(x.reshape(-1, x.shape[-1]))
            ws = [_nv.f32(getattr(seq[li], k)) for li in (0, 2, 4) for k in ("weight", "bias")]
            y = torch.empty((h.shape[0], seq[4].out_features), dtype=torch.float32, device=h.device)
            keep += [h] + ws
            p = probs[i]
            p.x, p.w0, p.b0, p.w2, p.b2, p.w4, p.b4 = [t.data_ptr() for t in [h] + ws]
            p.y, p.M, p.in_dim, p.hidden, p.out_dim = y.data_ptr(), h.shape[0], h.shape[1], 256, seq[4].out_features
            outs.append(y)
        _nv.check(_nv.lib().pstl_mlp3_batch(probs, len(jobs), _nv.stream()), "pstl_mlp3_batch")
        return outs

    def _mlp(self, seq, x):
        """encoder MLP.  Inference on the GPU goes through pstl_linear (row-independent summation order,
        so a scene's feature does not depend on which other scenes share its batch / shard); with autograd
        on (training, out of scope) it is the plain nn.Sequential."""
        if torch.is_grad_enabled() or not x.is_cuda:
            return seq(x)
        lead = x.shape[:-1]
        h = _nv.f32(x.reshape(-1, x.shape[-1]))
        L = _nv.lib()
        if len(seq) == 5 and seq[0].out_features == 256 and seq[2].out_features == 256 and seq[0].in_features <= 256 \
                and seq[4].out_features <= 32:
            ws = [_nv.f32(getattr(seq[li], k)) for li in (0, 2, 4) for k in ("weight", "bias")]
            y = torch.empty((h.shape[0], seq[4].out_features), dtype=torch.float32, device=h.device)
            _nv.check(L.pstl_mlp3(_nv.fptr(h), h.shape[0], h.shape[1], *[_nv.fptr(t) for t in ws], 256,
                                  seq[4].out_features, _nv.fptr(y), _nv.stream()), "pstl_mlp3")
            return y.reshape(*lead, -1)
        for li in (0, 2, 4):
            w, b = _nv.f32(seq[li].weight), _nv.f32(seq[li].bias)
            y = torch.empty((h.shape[0], w.shape[0]), dtype=torch.float32, device=h.device)
            _nv.check(L.pstl_linear(_nv.fptr(h), _nv.fptr(w), _nv.fptr(b), h.shape[0], h.shape[1], w.shape[0],
                                    int(li != 4), _nv.fptr(y), _nv.stream()), "pstl_linear")
            h = y
        return h.reshape(*lead, -1)

    # --- eps model (reference nusc_model.py:97-180, diffusion + multi_check branch) -------
    def forward(self, nn_input, ext=None, get_feature=False, prev_feature=None, sample=False, n_randoms=None):
        if getattr(self.args, "gt_data_training", False):
            raise NotImplementedError("--gt_data_training is out of scope")
        bs = nn_input["ego_traj"].shape[0]
        if n_randoms is None:
            n_randoms = self.args.n_randoms
        n_rep = n_randoms * 3
        scene_feat = getattr(prev_feature, "_pstl_scene_feat", None) if prev_feature is not None else None
        if scene_feat is None:
            if prev_feature is not None:
                scene_feat = prev_feature.reshape(bs, -1, prev_feature.shape[-1])[:, 0].contiguous()
            else:
                scene_feat = self.encode_feat(nn_input)
        x = _nv.f32(ext["noise"])
        _nv.require_cuda(x, "ext['noise']")
        n = x.shape[0]
        t = ext["timestep"]
        hl = _nv.f32(ext["highlevel"].reshape(n))
        stlp = _nv.f32(nn_input["stlp_dense"][:, 0])
        params = [p for li in (0, 2, 4) for p in (self.policy_net[li].weight, self.policy_net[li].bias)]
        if torch.is_grad_enabled() and (scene_feat.requires_grad or any(p.requires_grad for p in params)):
            # training (reference nusc_train.py:1352-1356): one timestep per row, gradients into policy_net and,
            # through the scene feature, into the encoders
            temb_rows = self.pos_encoding(t.reshape(n, 1).to(x.device), self.time_dim).to(torch.float32).contiguous()
            eps = _EpsRows.apply(self, scene_feat, hl.detach(), stlp.detach(), x.detach(), temb_rows, *params)
            controls = eps.reshape(-1, self.args.nt, 2)
            if get_feature:
                k = scene_feat.shape[-1]
                feature = scene_feat.reshape(bs, 1, k).expand(bs, n_rep, k).reshape(-1, k)
                feature._pstl_scene_feat = scene_feat
                return controls, feature
            return controls
        t0 = int(t.reshape(-1)[0].item())
        if not bool((t == t0).all()):
            raise NotImplementedError("native eps without autograd needs one timestep per call (the sampler's case)")
        temb_row = self.pos_encoding(torch.tensor([[t0]], dtype=torch.long), self.time_dim).to(x.device).contiguous()
        handle = self.native_handle("fp32")
        eps = torch.empty_like(x)
        L = _nv.lib()
        ws = _nv.workspace(L.pstl_denoiser_workspace_bytes(handle, n, bs, None), x.device, "denoiser")
        _nv.check(L.pstl_denoiser_eps(handle, _nv.fptr(_nv.f32(scene_feat)), bs, n // bs, _nv.fptr(hl), _nv.fptr(stlp),
                                      _nv.fptr(x), n, _nv.fptr(temb_row), _nv.fptr(eps), _nv.ptr(ws), _nv.stream()),
                  "pstl_denoiser_eps")
        controls = eps.reshape(-1, self.args.nt, 2)
        if get_feature:
            k = scene_feat.shape[-1]
            feature = scene_feat.reshape(bs, 1, k).expand(bs, n_rep, k).reshape(-1, k)
            feature._pstl_scene_feat = scene_feat
            return controls, feature
        return controls

    # --- RefineNet (reference nusc_model.py:182-235) --------------------------------------
    def _rect_native(self):
        """the fused RefineNet kernels (pstl_refine) cover the README configuration: --diverse_loss with the shard
        max-pool added to the controls, tanh-interval head"""
        a = self.args
        return bool(a.diverse_loss and not a.no_arch and a.diverse_fuse_type == "add" and a.interval)

    def _rect_forward_generic(self, feature, highlevel, stlp_dense_feat, init_controls, scores):
        """The other RefineNet variants of the reference (nusc_model.py:182-235): --no_arch / no --diverse_loss (no shard
        pooling), --diverse_fuse_type cat (pooled controls as extra inputs), no --interval (raw head).  Inference only:
        the dense layers run on pstl_linear (``_mlp``), the pooling / head arithmetic are upstream's tensor expressions."""
        a = self.args
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.rect_net.parameters()):
            raise NotImplementedError("training is built for the README RefineNet configuration only "
                                      "(--diverse_loss, --diverse_fuse_type add, --interval)")
        n, T2 = init_controls.shape[0], a.nt * 2
        sf = getattr(feature, "_pstl_scene_feat", None)
        if sf is not None:  # per-scene rows from this package's sampler
            feature = sf.reshape(sf.shape[0], 1, -1).expand(-1, n // sf.shape[0], -1).reshape(n, -1)
        init_controls = init_controls.reshape(n, a.nt, 2)
        if a.diverse_loss and not a.no_arch:
            fused = self._mlp(self.merge_net, init_controls.reshape(n, T2))
            bs, NS = n // 3 // a.n_randoms, a.n_shards
            fused = fused.reshape(bs, a.n_randoms, 3, T2).permute(0, 2, 1, 3).reshape(bs, 3, NS, a.n_randoms // NS, T2)
            fused = torch.max(fused, dim=3, keepdim=True)[0].repeat(1, 1, 1, a.n_randoms // NS, 1)
            fused = fused.reshape(bs, 3, a.n_randoms, T2).permute(0, 2, 1, 3).reshape(n, a.nt, 2)
            if a.diverse_fuse_type == "add":
                fused = init_controls + fused
                cols = [feature, highlevel, stlp_dense_feat, fused.reshape(n, T2)]
            elif a.diverse_fuse_type == "cat":
                cols = [feature, highlevel, stlp_dense_feat, init_controls.reshape(n, T2), fused.reshape(n, T2)]
            else:
                raise NotImplementedError("--diverse_fuse_type %s" % a.diverse_fuse_type)
        else:
            cols = [feature, highlevel, stlp_dense_feat, init_controls.reshape(n, T2)]
        raw = self._mlp(self.rect_net, torch.cat([c.reshape(n, -1) for c in cols], dim=-1)).reshape(n, a.nt, 2)
        if a.interval:
            raw = torch.tanh(raw)
            lim = torch.tensor([a.mul_w_max, a.mul_a_max], device=raw.device)
            mk = (raw >= 0).float()
            raw = (raw * (init_controls + lim)) * (1 - mk) + (raw * (lim - init_controls)) * mk
        out = init_controls + raw * (scores < 0).float()[:, None, None]
        if a.clip_rect:
            out = torch.stack([torch.clip(out[..., 0], -a.mul_w_max, a.mul_w_max),
                               torch.clip(out[..., 1], -a.mul_a_max, a.mul_a_max)], dim=-1)
        return out

    def rect_forward(self, feature, highlevel, stlp_dense_feat, init_controls, scores, extras=None):
        a = self.args
        if not self._rect_native():
            return self._rect_forward_generic(feature, highlevel, stlp_dense_feat, init_controls, scores)
        n = init_controls.shape[0]
        bs = int(n / 3 / a.n_randoms)
        scene_feat = getattr(feature, "_pstl_scene_feat", None)
        if scene_feat is None:
            scene_feat = feature.reshape(bs, -1, feature.shape[-1])[:, 0].contiguous()
        u0 = _nv.f32(init_controls.reshape(n, a.nt * 2))
        _nv.require_cuda(u0, "init_controls")
        params = [p for li in (0, 2, 4) for p in (self.rect_net[li].weight, self.rect_net[li].bias)]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            # training step (reference nusc_train.py:1228-1233: Adam over rect_net only): fp32 forward, native backward
            if getattr(a, "joint", False):
                raise NotImplementedError("--joint (gradients into merge_net / the encoders / the denoiser) is not built")
            return _RectForward.apply(self, _nv.f32(scene_feat), _nv.f32(highlevel.reshape(n)).detach(),
                                      _nv.f32(stlp_dense_feat.reshape(n, 6)).detach(), u0.detach(),
                                      _nv.f32(scores.reshape(n)).detach(), *params)
        out = torch.empty((n, a.nt, 2), dtype=torch.float32, device=u0.device)
        handle = self.native_handle(getattr(a, "precision", "fp32"))
        L = _nv.lib()
        ws = _nv.workspace(L.pstl_denoiser_workspace_bytes(handle, n, bs, None), u0.device, "denoiser")
        _nv.check(L.pstl_refine(handle, _nv.fptr(_nv.f32(scene_feat)), bs, n // bs, _nv.fptr(_nv.f32(highlevel.reshape(n))),
                                _nv.fptr(_nv.f32(stlp_dense_feat.reshape(n, 6))), _nv.fptr(u0),
                                _nv.fptr(_nv.f32(scores.reshape(n))), n, a.n_randoms, a.n_shards,
                                _nv.C.c_float(a.mul_w_max), _nv.C.c_float(a.mul_a_max), int(bool(a.clip_rect)),
                                _nv.fptr(out), _nv.ptr(ws), _nv.stream()), "pstl_refine")
        return out
