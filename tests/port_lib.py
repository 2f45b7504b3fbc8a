"""Loader for the C++ port of the reference algorithms (oracle/ecport.cpp) — checker / CPU baseline."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libecport.so")


def load():
    src = os.path.join(ROOT, "oracle", "ecport.cpp")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(LIB)
    lib.port_mul_gen.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
    lib.port_mul_var.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
    lib.port_verify.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p]
    return lib
