"""pstl_b200 — B200-native (sm_100a) hot path of pSTL-diffusion-policy behind the reference's Python API.

Sub-modules mirror the reference's flat files for the hot path only:
  stl_d_lib   – STL node classes (reference stl_d_lib.py)
  nusc_model  – Net (denoiser + RefineNet) (reference nusc_model.py)
  nusc_train  – rollout / STL spec / sampler / scoring / parser (reference nusc_train.py hot functions)
  native      – ctypes binding of the C-ABI library (include/pstl.h)
  synthetic   – seeded synthetic scenes (NuScenes is unavailable offline)
The CUDA library is loaded on first use; every op raises if it is missing (no CPU fallback).
"""
__version__ = "0.1.0"
