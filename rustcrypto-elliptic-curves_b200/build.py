"""Builds libecb200.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

    python -m importlib ... or: python rustcrypto-elliptic-curves_b200/build.py [--force]

One translation unit per curve + the host runtime, compiled in parallel, linked with -shared.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libecb200.so")
UNITS = ["abi", "curve_k256", "curve_p256", "curve_p384", "curve_sm2", "curve_p192", "curve_p224"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "ecb200.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(unit):
    src = os.path.join(CSRC, unit + ".cu")
    obj = os.path.join(OBJ, unit + ".o")
    log = os.path.join(OBJ, unit + ".ptxas.log")
    cmd = ["nvcc"] + NVCC_FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (unit, (r.stdout + r.stderr)[-4000:]))
    return obj


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(_compile, UNITS))
    cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
