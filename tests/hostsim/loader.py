"""Build + load the g++ host simulation of the per-trajectory device math (test tier only)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "hostsim.so")
SRC = os.path.join(HERE, "hostsim.cpp")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "pstl-diffusion-policy_b200", "csrc")


def load():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", SRC, "-o", SO])
    return C.CDLL(SO)
