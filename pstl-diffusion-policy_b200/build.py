"""In-tree build of libpstl_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python pstl-diffusion-policy_b200/build.py [--force]

The .so is written next to this file so it travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpstl_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# translation unit -> extra flags.  The STL/predicate units mirror the reference's separate
# fp32 roundings (no fused multiply-add contraction); the MLP units keep FMA.
UNITS = {
    "api.cu": [],
    "stl_kernels.cu": ["-fmad=false"],
    "aux_kernels.cu": ["-fmad=false"],
    "mlp_fp32.cu": [],
    "denoiser_tc.cu": [],
    "metrics.cu": [],
    "losses.cu": [],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(repr(sorted(UNITS.items())).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = os.path.join(HERE, "build", "stamp")
    units = dict(UNITS)
    if os.environ.get("PSTL_BUILD_TC_DEBUG"):
        # developer build: clock64 timeline stamps inside k_denoiser_tc (tests/tc_timeline.py); never shipped
        units["denoiser_tc.cu"] = units["denoiser_tc.cu"] + ["-DPSTL_TC_DEBUG"]
    dig = _digest() + ("+tcdebug" if os.environ.get("PSTL_BUILD_TC_DEBUG") else "")
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for unit, extra in units.items():
        obj = os.path.join(HERE, "build", unit.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, unit), "-o", obj]
        procs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for unit, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % unit)
    cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    open(stamp, "w").write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
