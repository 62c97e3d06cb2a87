"""ctypes binding of libpstl_b200.so (include/pstl.h).

There is no CPU fallback: every entry point raises ``PstlNativeError`` when the library is
missing or a call fails.  PyTorch is used only to own device memory and streams.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpstl_b200.so")

# opcodes / enums (mirror include/pstl.h)
OP_SIGNAL, OP_PRED, OP_NEG, OP_SMIN2, OP_SMAX2, OP_SMIN_K, OP_WIN_SMIN, OP_WIN_SMAX, OP_PREFIX_SMIN, OP_SUFFIX_SMAX = range(10)
SIG_V, SIG_D_CURR, SIG_TH_CURR, SIG_D_LEFT, SIG_TH_LEFT, SIG_D_RIGHT, SIG_TH_RIGHT, SIG_NEI = range(8)
DEN_ONE, DEN_THMAX, DEN_VFACTOR, DEN_DFACTOR, DEN_SFACTOR = range(5)
PRECISION_FP32, PRECISION_BF16, PRECISION_F16X3, PRECISION_F16 = 0, 1, 2, 3

EXPORTS = [
    "pstl_last_error", "pstl_device_info", "pstl_version", "pstl_program_create", "pstl_program_destroy",
    "pstl_program_tape_floats", "pstl_stl_workspace_bytes", "pstl_stl_eval_signals", "pstl_stl_eval_signals_bwd",
    "pstl_score_workspace_bytes", "pstl_score_fused", "pstl_score_fused_bwd", "pstl_guidance_step",
    "pstl_denoiser_create", "pstl_denoiser_destroy", "pstl_denoiser_workspace_bytes", "pstl_denoiser_sample",
    "pstl_denoiser_eps", "pstl_denoiser_set_noise_counter", "pstl_launch_count", "pstl_refine", "pstl_rollout", "pstl_rollout_bwd", "pstl_predicates", "pstl_linear", "pstl_encoder_inputs",
    "pstl_encoder_pool", "pstl_mlp3", "pstl_mlp3_batch", "pstl_trajopt_step", "pstl_diversity", "pstl_accuracy",
    "pstl_refine_losses_workspace_bytes", "pstl_refine_losses", "pstl_refine_backward_workspace_bytes",
    "pstl_refine_backward", "pstl_denoiser_eps_rows", "pstl_denoiser_eps_backward", "pstl_car_distances",
    "pstl_denoiser_refresh", "pstl_denoiser_set_engine",
]


class PstlNativeError(RuntimeError):
    pass


class Op(C.Structure):
    _fields_ = [("op", C.c_int32), ("a0", C.c_int32), ("a1", C.c_int32)]


class SceneView(C.Structure):
    _fields_ = [("neighbors", C.c_void_p), ("lanes", C.c_void_p * 3), ("n_scenes", C.c_int), ("Knei", C.c_int),
                ("nseg", C.c_int), ("T", C.c_int), ("rows_per_scene", C.c_int)]


class SpecParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("tau", C.c_float), ("ego_L", C.c_float), ("ego_W", C.c_float),
                ("w_scale", C.c_float), ("a_scale", C.c_float), ("clip_controls", C.c_int), ("clip_dist", C.c_int),
                ("hard", C.c_int)]


_W_FIELDS = ["p0_w", "p0_b", "p2_w", "p2_b", "p4_w", "p4_b", "m0_w", "m0_b", "m2_w", "m2_b", "m4_w", "m4_b",
             "r0_w", "r0_b", "r2_w", "r2_b", "r4_w", "r4_b"]


class Weights(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in _W_FIELDS] + [(f, C.c_int) for f in
                                                        ("hidden", "rect_hidden", "merge_hidden", "feat_dim", "time_dim", "T")]


class Mlp3Problem(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in ("x", "w0", "b0", "w2", "b2", "w4", "b4", "y")] + \
               [(f, C.c_int) for f in ("M", "in_dim", "hidden", "out_dim")]


class LossCfg(C.Structure):
    _fields_ = [(f, C.c_int) for f in ("n_scenes", "S", "nt", "n_shards", "diverse_loss", "diverse_detach")] + \
               [(f, C.c_float) for f in ("w_max", "a_max", "stl_nn_thres", "stl_weight", "diversity_scale",
                                         "diversity_weight", "rect_reg_loss", "extra_rect_reg")]


LOSS_KEYS = ("loss", "loss_stl", "loss_reg", "loss_diversity", "extra_loss_reg")


class GuidanceCfg(C.Structure):
    _fields_ = [("valid", C.c_void_p), ("state0", C.c_void_p), ("progs", C.POINTER(C.c_void_p)),
                ("scenes", C.POINTER(SceneView)), ("sp", C.POINTER(SpecParams)), ("before", C.c_int),
                ("step_mask", C.c_void_p), ("niters", C.c_int), ("lr", C.c_float), ("thres", C.c_float),
                ("inv_norm", C.c_float), ("inv_norm_dev", C.c_void_p)]


_lib = None


def lib():
    """Load the shared library once; fail loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PstlNativeError(
            "libpstl_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python pstl-diffusion-policy_b200/build.py`. There is no CPU fallback." % LIB_PATH)
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise PstlNativeError("cannot load %s: %s" % (LIB_PATH, e))
    L.pstl_last_error.restype = C.c_char_p
    L.pstl_launch_count.restype = C.c_ulonglong
    for name in ("pstl_stl_workspace_bytes", "pstl_score_workspace_bytes", "pstl_denoiser_workspace_bytes",
                 "pstl_refine_losses_workspace_bytes", "pstl_refine_backward_workspace_bytes"):
        getattr(L, name).restype = C.c_size_t
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise PstlNativeError("%s failed (%d): %s" % (what, rc, lib().pstl_last_error().decode()))


def require_cuda(t, name="tensor"):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise PstlNativeError("%s must be a CUDA tensor: pstl_b200 has no CPU path" % name)


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def fptr(t, name="tensor"):
    """Pointer of a contiguous fp32 CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    require_cuda(t, name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise PstlNativeError("%s must be contiguous float32" % name)
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32(t):
    """contiguous fp32 view/copy of t on its device"""
    return t.detach().to(torch.float32).contiguous()


class Program:
    """Owns a pstl_program_t."""

    def __init__(self, ops, n_signals, T, need_t):
        arr = (Op * len(ops))(*[Op(int(o[0]), int(o[1]), int(o[2])) for o in ops])
        h = C.c_void_p()
        check(lib().pstl_program_create(arr, len(ops), int(n_signals), int(T), int(need_t), C.byref(h)),
              "pstl_program_create")
        self.h, self.ops, self.n_signals, self.T, self.need_t = h, tuple(ops), n_signals, T, need_t

    def __del__(self):
        try:
            if getattr(self, "h", None) and _lib is not None:
                _lib.pstl_program_destroy(self.h)
        except Exception:
            pass


_ws_cache = {}


def workspace(nbytes, device, tag="default"):
    """Grow-only scratch per (device, tag, stream); returned tensor is uint8."""
    if nbytes <= 0:
        return None
    key = (device, tag, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.1) + 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


_act_epoch = [0]


def activation_workspace(nbytes, device):
    """Scratch of a training forward whose activations the backward may reuse: returns (buffer, epoch); the epoch
    moves on with every later training forward on this buffer (``activations_valid``)."""
    _act_epoch[0] += 1
    return workspace(nbytes, device, "train_act"), _act_epoch[0]


def activations_valid(epoch):
    return int(_act_epoch[0] == epoch)


def make_scene_view(neighbors, lanes, rows_per_scene):
    """neighbors (n_scenes,K,T,7); lanes: 3 tensors (n_scenes,nseg,3)."""
    sv = SceneView()
    sv.neighbors = neighbors.data_ptr()
    for i in range(3):
        sv.lanes[i] = lanes[i].data_ptr()
    sv.n_scenes, sv.Knei, sv.T = neighbors.shape[0], neighbors.shape[1], neighbors.shape[2]
    sv.nseg = lanes[0].shape[1]
    sv.rows_per_scene = int(rows_per_scene)
    return sv


def make_spec(dt, tau, ego_L, ego_W, w_scale=1.0, a_scale=1.0, clip_controls=0, clip_dist=0, hard=0):
    return SpecParams(dt, tau, ego_L, ego_W, w_scale, a_scale, int(clip_controls), int(clip_dist), int(hard))


def prog_array(progs):
    return (C.c_void_p * 3)(*[p.h for p in progs])
