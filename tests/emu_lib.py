"""Loader for the TEST-ONLY host emulation of the kernel bodies (tests/emu/emu.cpp)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu.cpp")
LIB = os.path.join(HERE, "emu", "libecb200_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "rustcrypto-elliptic-curves_b200", "csrc")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    if _stale():
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", LIB, SRC])
    lib = ctypes.CDLL(LIB)
    return lib


def load_variant(kt_w):
    """the same emulation built with another window width of the per-key tables (-DECB_KT_W): every width the product may be
    built with is checked on the CPU tier"""
    path = LIB.replace(".so", "_ktw%d.so" % kt_w)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(path) or any(os.path.getmtime(d) > os.path.getmtime(path) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-DECB_KT_W=%d" % kt_w, "-o", path, SRC])
    return ctypes.CDLL(path)


def buf(n):
    return (ctypes.c_uint8 * n)()
