"""Second, independent CPU oracle and CPU figure: OpenSSL libcrypto through ctypes.

TEST / BENCH INFRASTRUCTURE ONLY — like everything under oracle/, it is imported by tests/, bench.py's cpu_baseline leg
and nothing else; the product never touches it.  It exists because the reference (Rust) cannot be built here: a
production-grade implementation that shares no code with the Python oracle or the C++ port pins the curve arithmetic a
second time (SM2 in particular, for which the reference holds a single k*G vector) and gives an honest "good CPU
library on the same cores" number next to the C++ port of the reference algorithms.

Semantics reproduced on top of libcrypto so that results are comparable with the reference's:
  * scalars are reduced once mod n before EC_POINT_mul (Reduce<Uint>::reduce_bytes);
  * ECDSA_do_verify + the explicit low-s rule for secp256k1 (k256/src/ecdsa.rs:203-205); keys that fail
    EC_POINT_set_affine_coordinates (off curve / >= p) and r, s outside [1, n-1] count as rejected.
"""
import ctypes
import ctypes.util
import os
from concurrent.futures import ThreadPoolExecutor

from . import ecoracle as o

NID = {"k256": 714, "p256": 415, "p384": 715, "sm2": 1172, "p192": 409, "p224": 713}   # 409 = NID_X9_62_prime192v1, 713 = NID_secp224r1
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    name = ctypes.util.find_library("crypto")
    if not name:
        raise RuntimeError("libcrypto not found")
    L = ctypes.CDLL(name)
    vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    for fn, res, args in (
        ("EC_GROUP_new_by_curve_name", vp, [ci]), ("EC_POINT_new", vp, [vp]), ("EC_POINT_free", None, [vp]),
        ("BN_bin2bn", vp, [ctypes.c_char_p, ci, vp]), ("BN_free", None, [vp]), ("BN_CTX_new", vp, []), ("BN_CTX_free", None, [vp]),
        ("EC_POINT_mul", ci, [vp, vp, vp, vp, vp, vp]), ("EC_POINT_set_affine_coordinates", ci, [vp, vp, vp, vp, vp]),
        ("EC_POINT_point2oct", sz, [vp, vp, ci, ctypes.c_char_p, sz, vp]), ("EC_POINT_is_at_infinity", ci, [vp, vp]),
        ("EC_KEY_new", vp, []), ("EC_KEY_free", None, [vp]), ("EC_KEY_set_group", ci, [vp, vp]), ("EC_KEY_set_public_key", ci, [vp, vp]),
        ("ECDSA_SIG_new", vp, []), ("ECDSA_SIG_free", None, [vp]), ("ECDSA_SIG_set0", ci, [vp, vp, vp]),
        ("ECDSA_do_verify", ci, [ctypes.c_char_p, ci, vp, vp]), ("ERR_clear_error", None, []),
    ):
        f = getattr(L, fn)
        f.restype, f.argtypes = res, args
    L.OpenSSL_version.restype, L.OpenSSL_version.argtypes = ctypes.c_char_p, [ci]
    _lib = L
    return L


_groups = {}


def group(cname):
    if cname not in _groups:
        g = lib().EC_GROUP_new_by_curve_name(NID[cname])
        if not g:
            raise RuntimeError("libcrypto has no curve %s" % cname)
        _groups[cname] = g
    return _groups[cname]


def _chunks(n, parts):
    step = (n + parts - 1) // parts
    return [(i, min(n, i + step)) for i in range(0, n, step)]


def _pool_map(fn, n, threads):
    threads = threads or os.cpu_count() or 1
    if n == 0:
        return b""
    with ThreadPoolExecutor(max_workers=threads) as ex:      # ctypes releases the GIL inside libcrypto calls
        return b"".join(ex.map(lambda r: fn(*r), _chunks(n, threads * 4)))


_drv = None


def driver():
    """oracle/libosslref.so (osslref.c: the same calls from C with OpenMP, no per-row Python overhead) or None."""
    global _drv
    if _drv is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "libosslref.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.call(["make", "-s", "-C", here], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        try:
            d = ctypes.CDLL(path)
            d.ossl_mul.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p]
            d.ossl_verify.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int,
                                      ctypes.c_char_p]
            d.ossl_set_threads.argtypes = [ctypes.c_int]
            _drv = d
        except OSError:
            _drv = False
    return _drv or None


def threads_used(threads=None):
    d = driver()
    if d:
        d.ossl_set_threads(threads or os.cpu_count() or 1)
        return d.ossl_threads()
    return threads or os.cpu_count() or 1


def mul_batch(cname, ks: bytes, pts_xy: bytes = None, compress=None, threads=None) -> bytes:
    """SEC1 slots of k_i * G (pts_xy None) or k_i * P_i (P_i = x||y affine)."""
    L, c, g = lib(), o.curve(cname), group(cname)
    fb = c.fb
    compress = c.compress if compress is None else compress
    slot = 1 + (fb if compress else 2 * fb)
    n = len(ks) // fb
    d = driver()
    if d:
        d.ossl_set_threads(threads or os.cpu_count() or 1)
        out = ctypes.create_string_buffer(max(1, n * slot))
        if d.ossl_mul(NID[cname], fb, n, pts_xy, ks, 1 if compress else 0, out) != 0:
            raise RuntimeError("ossl_mul failed")
        return out.raw[:n * slot]

    def work(lo, hi):
        ctx = L.BN_CTX_new()
        P, R = L.EC_POINT_new(g), L.EC_POINT_new(g)
        out = bytearray()
        buf = ctypes.create_string_buffer(slot)
        for i in range(lo, hi):
            k = o.reduce_once(c, int.from_bytes(ks[fb * i:fb * i + fb], "big"))
            bk = L.BN_bin2bn(k.to_bytes(fb, "big"), fb, None)
            if pts_xy is None:
                ok = L.EC_POINT_mul(g, R, bk, None, None, ctx)
            else:
                bx = L.BN_bin2bn(pts_xy[2 * fb * i:2 * fb * i + fb], fb, None)
                by = L.BN_bin2bn(pts_xy[2 * fb * i + fb:2 * fb * i + 2 * fb], fb, None)
                ok = L.EC_POINT_set_affine_coordinates(g, P, bx, by, ctx) and L.EC_POINT_mul(g, R, None, P, bk, ctx)
                L.BN_free(bx); L.BN_free(by)
            L.BN_free(bk)
            if not ok or L.EC_POINT_is_at_infinity(g, R):
                L.ERR_clear_error()
                out += bytes(slot)
                continue
            m = L.EC_POINT_point2oct(g, R, 2 if compress else 4, buf, slot, ctx)
            out += buf.raw[:m] + bytes(slot - m)
        L.EC_POINT_free(P); L.EC_POINT_free(R); L.BN_CTX_free(ctx)
        return bytes(out)

    return _pool_map(work, n, threads)


def verify_batch(cname, q: bytes, z: bytes, rs: bytes, threads=None) -> bytes:
    """ok bytes with verify_prehashed semantics (incl. the k256 low-s rule)."""
    L, c, g = lib(), o.curve(cname), group(cname)
    fb = c.fb
    n = len(z) // fb
    d = driver()
    if d:
        d.ossl_set_threads(threads or os.cpu_count() or 1)
        out = ctypes.create_string_buffer(max(1, n))
        if d.ossl_verify(NID[cname], fb, n, q, z, rs, 1 if c.low_s else 0, out) != 0:
            raise RuntimeError("ossl_verify failed")
        return out.raw[:n]

    def work(lo, hi):
        ctx = L.BN_CTX_new()
        P = L.EC_POINT_new(g)
        out = bytearray()
        for i in range(lo, hi):
            r = int.from_bytes(rs[2 * fb * i:2 * fb * i + fb], "big")
            s = int.from_bytes(rs[2 * fb * i + fb:2 * fb * i + 2 * fb], "big")
            if not (1 <= r < c.n and 1 <= s < c.n) or (c.low_s and s > c.n >> 1):
                out.append(0)
                continue
            bx = L.BN_bin2bn(q[2 * fb * i:2 * fb * i + fb], fb, None)
            by = L.BN_bin2bn(q[2 * fb * i + fb:2 * fb * i + 2 * fb], fb, None)
            ok = L.EC_POINT_set_affine_coordinates(g, P, bx, by, ctx)
            L.BN_free(bx); L.BN_free(by)
            if not ok:
                L.ERR_clear_error()
                out.append(0)
                continue
            key = L.EC_KEY_new()
            L.EC_KEY_set_group(key, g)
            L.EC_KEY_set_public_key(key, P)
            sig = L.ECDSA_SIG_new()
            L.ECDSA_SIG_set0(sig, L.BN_bin2bn(rs[2 * fb * i:2 * fb * i + fb], fb, None), L.BN_bin2bn(rs[2 * fb * i + fb:2 * fb * i + 2 * fb], fb, None))
            v = L.ECDSA_do_verify(z[fb * i:fb * i + fb], fb, sig, key)
            if v != 1:
                L.ERR_clear_error()
            out.append(1 if v == 1 else 0)
            L.ECDSA_SIG_free(sig); L.EC_KEY_free(key)
        L.EC_POINT_free(P); L.BN_CTX_free(ctx)
        return bytes(out)

    return _pool_map(work, n, threads)
