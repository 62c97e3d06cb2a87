"""Import the UNMODIFIED reference from /root/reference in the build container (test tooling only).

The reference needs matplotlib, imageio and the nuscenes devkit at import time
(utils.py:12-13, nusc_api.py:1-11, nusc_viz.py:2-4); none are installed here, so
empty stub modules are registered first.  On a
CPU-only host ``.cuda()`` is patched to identity.  /root/reference does not exist
on the GPU box: only make_golden.py and the ``ref``-marked CPU tests use this file.
"""
import os
import sys
import types

REF_DIR = os.environ.get("PSTL_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "stl_d_lib.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


class _Dummy:
    def __init__(self, *a, **k):
        pass


def install_stubs():
    import torch
    plt = _stub("matplotlib.pyplot")
    _stub("matplotlib", pyplot=plt)
    _stub("matplotlib.patches", Polygon=_Dummy, Rectangle=_Dummy, Ellipse=_Dummy, Circle=_Dummy)
    _stub("matplotlib.ticker", PercentFormatter=_Dummy)
    _stub("imageio")
    _stub("nuscenes")
    _stub("nuscenes.nuscenes", NuScenes=_Dummy, NuScenesExplorer=_Dummy)
    _stub("nuscenes.map_expansion")
    _stub("nuscenes.map_expansion.map_api", NuScenesMap=_Dummy)
    _stub("nuscenes.map_expansion.arcline_path_utils")
    _stub("nuscenes.utils")
    _stub("nuscenes.utils.map_mask", MapMask=_Dummy)
    _stub("nuscenes.utils.color_map", get_colormap=lambda *a, **k: {})
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def load(argv=None):
    """Return (nusc_train module, parsed args).  ``argv`` = README-style flag list."""
    assert available(), "reference not present"
    install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    old = sys.argv
    sys.argv = ["nusc_train.py"] + list(argv or [])
    try:
        import nusc_train as T
        args = T.generate_parser()
    finally:
        sys.argv = old
    T.args = args
    T.plot_debug_scene = lambda *a, **k: None
    return T, args


OURS_FLAGS = ["-e", "e7_ours", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--rect_head", "--flex",
              "--diverse_loss", "--multi_cands", "5", "--test", "--run_sampling_test", "--skip_nusc_load",
              "--viz_correct"]
GUIDE_FLAGS = ["-e", "e7_ours", "--diffusion", "--stl_weight", "0.0", "--load_stlp", "--rect_head", "--flex",
               "--diverse_loss", "--multi_cands", "10", "--test", "--run_sampling_test", "--viz_correct",
               "--guidance", "--guidance_before", "10", "--guidance_niters", "1", "--guidance_lr", "0.01",
               "--n_rolls", "3", "--other", "--skip_nusc_load"]
