"""ecb200 — host-side mirror of the reference's batch-able trait surface over the C ABI.

The reference is Rust (no toolchain in this image), so this module plays the role of the wrapper
crate: same names, argument meaning and error behaviour as the reference operators on the hot
path, each a thin ctypes call into ``libecb200.so`` (include/ecb200.h).  There is NO CPU fallback:
if the CUDA library is missing or no B200 is visible, construction raises.

    reference item                                              here
    ----------------------------------------------------------  ---------------------------------------
    ProjectivePoint::mul_by_generator(&k)  (k256 mul.rs:415-440) Engine.mul_by_generator_batch(curve, ks)
    &P * &k  + batch_normalize              (mul.rs:443-481)     Engine.mul_batch(curve, points, ks)
    BatchNormalize::batch_normalize         (projective.rs:325)  Engine.batch_normalize(curve, xyz)
    LinearCombinationExt::lincomb_ext       (mul.rs:313-340)     Engine.lincomb(curve, points, ks)
    VerifyingKey::verify_prehash            (ecdsa.rs:200-209)   Engine.verify_prehash_batch(curve, keys, prehashes, sigs)
    FieldElement / Scalar ops (test hook)   (field_8x32_risc0.rs) Engine.field_op(curve, which, op, a, b)
    AffinePoint::from_encoded_point / decompress (affine.rs:184-270) Engine.decode_points(curve, enc, stride, mode)
    VerifyingKey::from_sec1_bytes + verify_prehash               Engine.ecdsa_verify_sec1(curve, keys, stride, z, rs)
    VerifyingKey::recover_from_prehash      (ecdsa.rs:278-343)   Engine.ecdsa_recover(curve, z, rs, recid)
    schnorr::VerifyingKey::verify_prehash   (schnorr/verifying.rs:63-89) Engine.schnorr_verify(pk, e, sig)
    sm2::dsa::VerifyingKey::verify_prehash  (sm2 dsa/verifying.rs:130-168) Engine.sm2dsa_verify(q, e, rs)
    SignPrimitive::try_sign_prehashed       (ecdsa.rs:181-198)   Engine.ecdsa_sign(curve, d, k, z)
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, List, Optional, Sequence, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ECB200_LIB") or os.path.join(HERE, "libecb200.so")

K256, P256, P384, SM2, P192, P224 = 0, 1, 2, 3, 4, 5
CURVE_IDS = {"k256": K256, "secp256k1": K256, "p256": P256, "p384": P384, "sm2": SM2, "p192": P192, "p224": P224}

FLAG_CT = 1
FLAG_COMPRESSED = 2
FLAG_UNCOMPRESSED = 4
FLAG_PROJ = 8

# every symbol include/ecb200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "ecb200_field_bytes", "ecb200_point_slot_bytes", "ecb200_init", "ecb200_destroy", "ecb200_last_error",
    "ecb200_version", "ecb200_launch_count", "ecb200_sync", "ecb200_mul_gen", "ecb200_mul_var",
    "ecb200_batch_normalize", "ecb200_lincomb", "ecb200_ecdsa_verify", "ecb200_field_op", "ecb200_mul_gen_dev",
    "ecb200_mul_var_dev", "ecb200_batch_normalize_dev", "ecb200_ecdsa_verify_dev",
    "ecb200_decode_points", "ecb200_ecdsa_verify_sec1", "ecb200_ecdsa_recover", "ecb200_schnorr_verify", "ecb200_sm2dsa_verify",
    "ecb200_ecdsa_sign", "ecb200_decode_points_dev", "ecb200_ecdsa_verify_sec1_dev", "ecb200_ecdsa_recover_dev",
    "ecb200_schnorr_verify_dev", "ecb200_sm2dsa_verify_dev", "ecb200_ecdsa_sign_dev",
    "ecb200_kernel_timing", "ecb200_kernel_timing_read",
    "ecb200_init_multi", "ecb200_device_count", "ecb200_lincomb2", "ecb200_lincomb2_dev", "ecb200_keytab_stats",
]
DECODE_SEC1, DECODE_COMPACT = 0, 1


class Ecb200Error(RuntimeError):
    pass


_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen libecb200.so and declare prototypes.  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Ecb200Error("libecb200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                          "there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    u8p, vp, sz, u32, ci = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int
    lib.ecb200_field_bytes.restype = sz
    lib.ecb200_field_bytes.argtypes = [ci]
    lib.ecb200_point_slot_bytes.restype = sz
    lib.ecb200_point_slot_bytes.argtypes = [ci, u32]
    lib.ecb200_init.argtypes = [ci, ctypes.POINTER(vp)]
    lib.ecb200_init_multi.argtypes = [ci, ctypes.POINTER(ci), ctypes.POINTER(vp)]
    lib.ecb200_device_count.argtypes = [vp]
    lib.ecb200_lincomb2.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p, u8p, u8p, u32]
    lib.ecb200_lincomb2_dev.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p, u8p, u8p, u32, vp]
    lib.ecb200_destroy.argtypes = [vp]
    lib.ecb200_destroy.restype = None
    lib.ecb200_last_error.argtypes = [vp]
    lib.ecb200_last_error.restype = ctypes.c_char_p
    lib.ecb200_version.restype = ctypes.c_char_p
    lib.ecb200_launch_count.argtypes = [vp]
    lib.ecb200_launch_count.restype = ctypes.c_uint64
    lib.ecb200_sync.argtypes = [vp]
    lib.ecb200_keytab_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    lib.ecb200_kernel_timing.argtypes = [vp, ci]
    lib.ecb200_kernel_timing_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    lib.ecb200_mul_gen.argtypes = [vp, ci, sz, u8p, u8p, u32]
    lib.ecb200_mul_var.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p, u8p, u32]
    lib.ecb200_batch_normalize.argtypes = [vp, ci, sz, u8p, u8p, u8p]
    lib.ecb200_lincomb.argtypes = [vp, ci, sz, u8p, u8p, u8p, u32, u32]
    lib.ecb200_ecdsa_verify.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p]
    lib.ecb200_field_op.argtypes = [vp, ci, ci, ci, sz, u8p, u8p, u8p, u8p]
    lib.ecb200_mul_gen_dev.argtypes = [vp, ci, sz, u8p, u8p, u32, vp]
    lib.ecb200_mul_var_dev.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p, u8p, u32, vp]
    lib.ecb200_batch_normalize_dev.argtypes = [vp, ci, sz, u8p, u8p, u8p, vp]
    lib.ecb200_ecdsa_verify_dev.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p, vp]
    lib.ecb200_decode_points.argtypes = [vp, ci, sz, u8p, sz, u32, u8p, u8p]
    lib.ecb200_ecdsa_verify_sec1.argtypes = [vp, ci, sz, u8p, sz, u8p, u8p, u8p]
    lib.ecb200_ecdsa_recover.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p, u8p, u32]
    lib.ecb200_schnorr_verify.argtypes = [vp, sz, u8p, u8p, u8p, u8p]
    lib.ecb200_sm2dsa_verify.argtypes = [vp, sz, u8p, u8p, u8p, u8p]
    lib.ecb200_ecdsa_sign.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p, u8p, u8p]
    lib.ecb200_decode_points_dev.argtypes = [vp, ci, sz, u8p, sz, u32, u8p, u8p, vp]
    lib.ecb200_ecdsa_verify_sec1_dev.argtypes = [vp, ci, sz, u8p, sz, u8p, u8p, u8p, vp]
    lib.ecb200_ecdsa_recover_dev.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p, u8p, u32, vp]
    lib.ecb200_schnorr_verify_dev.argtypes = [vp, sz, u8p, u8p, u8p, u8p, vp]
    lib.ecb200_sm2dsa_verify_dev.argtypes = [vp, sz, u8p, u8p, u8p, u8p, vp]
    lib.ecb200_ecdsa_sign_dev.argtypes = [vp, ci, sz, u8p, u8p, u8p, u8p, u8p, u8p, vp]
    _lib = lib
    return lib


def curve_id(curve) -> int:
    return CURVE_IDS[curve] if isinstance(curve, str) else int(curve)


def field_bytes(curve) -> int:
    cid = curve_id(curve)
    return 48 if cid == P384 else 24 if cid == P192 else 28 if cid == P224 else 32


def slot_bytes(curve, flags: int = 0) -> int:
    cid = curve_id(curve)
    comp = True if flags & FLAG_COMPRESSED else False if flags & FLAG_UNCOMPRESSED else cid == K256
    return 1 + (field_bytes(cid) if comp else 2 * field_bytes(cid))


def bits2field(curve, prehash: bytes) -> bytes:
    """ecdsa::hazmat::bits2field: < FB/2 bytes -> error; shorter -> left-pad;
    longer -> leftmost FB bytes.  Host-side, as in the reference (it is byte shuffling, not arithmetic)."""
    fb = field_bytes(curve)
    if len(prehash) < fb // 2:
        raise ValueError("prehash too short")            # signature::Error in the reference
    if len(prehash) < fb:
        return b"\x00" * (fb - len(prehash)) + prehash
    return prehash[:fb]


def _nbytes(b) -> int:
    return int(b.nbytes) if hasattr(b, "nbytes") else len(b)


def _as_buf(b, need: Optional[int] = None, what: str = "buffer") -> Tuple[ctypes.c_void_p, object]:
    """bytes / bytearray / numpy uint8 array -> (void*, keep-alive).  The C ABI takes raw pointers and trusts the row
    count, so the length is checked HERE: `need` is the exact byte size the call will read or write through the pointer
    (a shorter buffer would be read past its end).  numpy arrays must be uint8 and C-contiguous - a strided view such as
    a[:, :fb] would silently feed the wrong bytes."""
    if b is None:
        return None, None
    if isinstance(b, bytes):       # zero-copy view of the immutable buffer (inputs only)
        ptr, keep = ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), b
    elif isinstance(b, bytearray):
        keep = (ctypes.c_uint8 * len(b)).from_buffer(b)
        ptr = ctypes.cast(keep, ctypes.c_void_p)
    elif hasattr(b, "ctypes") and hasattr(b, "flags"):      # numpy
        if str(b.dtype) != "uint8":
            raise ValueError("%s: numpy buffers must have dtype uint8, got %s" % (what, b.dtype))
        if not b.flags.c_contiguous:
            raise ValueError("%s: numpy buffers must be C-contiguous (use np.ascontiguousarray)" % what)
        ptr, keep = ctypes.c_void_p(b.ctypes.data), b
    else:
        raise TypeError("%s: unsupported buffer type %r" % (what, type(b)))
    if need is not None and _nbytes(b) != need:
        raise ValueError("%s: expected %d bytes, got %d" % (what, need, _nbytes(b)))
    return ptr, keep


def _rows(b, elem: int, what: str) -> int:
    """row count of the buffer that defines n; its size must be a whole number of rows"""
    nb = _nbytes(b)
    if elem <= 0 or nb % elem:
        raise ValueError("%s: %d bytes is not a whole number of %d-byte rows" % (what, nb, elem))
    return nb // elem


class Engine:
    """One context.  Engine(device) = one GPU (the one-process-per-GPU deployment); Engine(devices=[0, 1, ...]) or
    Engine(devices="all") = one context over several GPUs of the box (ecb200_init_multi): the host-buffer methods shard
    every batch by contiguous index range over the devices inside ONE call.  Thread-compatible, not thread-safe."""

    def __init__(self, device: int = 0, devices=None):
        self.lib = load_library()
        h = ctypes.c_void_p()
        if devices is None:
            rc = self.lib.ecb200_init(int(device), ctypes.byref(h))
            what = "ecb200_init(device=%d)" % device
        elif isinstance(devices, str):
            if devices != "all":
                raise ValueError("devices must be a list of device ordinals or 'all'")
            rc = self.lib.ecb200_init_multi(0, None, ctypes.byref(h))
            what = "ecb200_init_multi(all devices)"
        else:
            devs = [int(d) for d in devices]
            arr = (ctypes.c_int * len(devs))(*devs)
            rc = self.lib.ecb200_init_multi(len(devs), arr, ctypes.byref(h))
            what = "ecb200_init_multi(%r)" % (devs,)
        if rc != 0 or not h:
            raise Ecb200Error("%s failed with %d: no usable CUDA device or kernel build mismatch "
                              "(there is no CPU fallback)" % (what, rc))
        self.h = h
        self.device = device
        self.n_devices = int(self.lib.ecb200_device_count(h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.ecb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise Ecb200Error("%s failed (%d): %s" % (what, rc, self.lib.ecb200_last_error(self.h).decode()))

    @property
    def launch_count(self) -> int:
        return int(self.lib.ecb200_launch_count(self.h))

    def sync(self):
        self._check(self.lib.ecb200_sync(self.h), "sync")

    def keytab_stats(self) -> Tuple[int, int]:
        """(rows verified on per-key window tables, tables built) since the engine was created."""
        rows, tabs = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self.lib.ecb200_keytab_stats(self.h, ctypes.byref(rows), ctypes.byref(tabs)), "keytab_stats")
        return int(rows.value), int(tabs.value)

    def kernel_timing(self, enable: bool):
        self._check(self.lib.ecb200_kernel_timing(self.h, 1 if enable else 0), "kernel_timing")

    def kernel_timing_read(self) -> Tuple[float, int]:
        """(summed ms, launches) of the dominant kernel since the last read (CUDA events on its launch stream)."""
        ms, cnt = ctypes.c_double(), ctypes.c_uint64()
        self._check(self.lib.ecb200_kernel_timing_read(self.h, ctypes.byref(ms), ctypes.byref(cnt)), "kernel_timing_read")
        return float(ms.value), int(cnt.value)

    # ------------------------------------------------------------------ host-buffer API (bytes in, bytes out)
    # Every method derives n from ONE buffer and checks every other buffer against it (ValueError on mismatch):
    # the C ABI takes raw pointers and would read past a short buffer.
    def mul_by_generator_batch(self, curve, ks: bytes, flags: int = 0) -> bytes:
        """[k_i * G] as SEC1 slots.  ks = n x FB big-endian scalars."""
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(ks, fb, "mul_gen: ks")
        out = bytearray(n * slot_bytes(cid, flags))
        pk, k1 = _as_buf(ks, n * fb, "mul_gen: ks")
        po, k2 = _as_buf(out)
        self._check(self.lib.ecb200_mul_gen(self.h, cid, n, pk, po, flags), "mul_gen")
        return bytes(out)

    def mul_batch(self, curve, points: bytes, ks: bytes, inf: Optional[bytes] = None, flags: int = 0) -> Tuple[bytes, bytes]:
        """[k_i * P_i] as SEC1 slots, plus the per-element invalid-point flags."""
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(ks, fb, "mul_var: ks")
        out = bytearray(n * slot_bytes(cid, flags))
        invalid = bytearray(n)
        pp, a = _as_buf(points, n * fb * (3 if flags & FLAG_PROJ else 2), "mul_var: points")
        pi, b = _as_buf(inf, n, "mul_var: inf")
        pk, c = _as_buf(ks, n * fb, "mul_var: ks")
        po, d = _as_buf(out)
        pv, e = _as_buf(invalid)
        self._check(self.lib.ecb200_mul_var(self.h, cid, n, pp, pi, pk, po, pv, flags), "mul_var")
        return bytes(out), bytes(invalid)

    def batch_normalize(self, curve, xyz: bytes) -> Tuple[bytes, bytes]:
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(xyz, 3 * fb, "batch_normalize: xyz")
        xy = bytearray(n * 2 * fb)
        inf = bytearray(n)
        pi, a = _as_buf(xyz, n * 3 * fb, "batch_normalize: xyz")
        po, b = _as_buf(xy)
        pf, c = _as_buf(inf)
        self._check(self.lib.ecb200_batch_normalize(self.h, cid, n, pi, po, pf), "batch_normalize")
        return bytes(xy), bytes(inf)

    def lincomb(self, curve, points: bytes, ks: bytes, flags: int = 0, out_proj: bool = False) -> bytes:
        """sum_i k_i * P_i as one SEC1 slot (or X||Y||Z when out_proj)."""
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(ks, fb, "lincomb: ks")
        out = bytearray(3 * fb if out_proj else slot_bytes(cid, flags))
        pp, a = _as_buf(points, n * fb * (3 if flags & FLAG_PROJ else 2), "lincomb: points")
        pk, b = _as_buf(ks, n * fb, "lincomb: ks")
        po, c = _as_buf(out)
        self._check(self.lib.ecb200_lincomb(self.h, cid, n, pp, pk, po, flags, FLAG_PROJ if out_proj else 0), "lincomb")
        return bytes(out)

    def lincomb2_batch(self, curve, p1: bytes, k1: bytes, p2: bytes, k2: bytes, flags: int = 0) -> Tuple[bytes, bytes]:
        """[k1_i * P1_i + k2_i * P2_i] as SEC1 slots, one result PER ROW, plus per-row invalid flags:
        LinearCombination::lincomb(&x, &k, &y, &l) (k256 mul.rs:313-323, primeorder projective.rs:415-420) over slices.
        Points are affine x||y (X||Y||Z with FLAG_PROJ); FLAG_CT selects the secret-scalar kernels."""
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(k1, fb, "lincomb2: k1")
        pb = fb * (3 if flags & FLAG_PROJ else 2)
        out = bytearray(n * slot_bytes(cid, flags))
        invalid = bytearray(n)
        a1, h1 = _as_buf(p1, n * pb, "lincomb2: p1")
        a2, h2 = _as_buf(k1, n * fb, "lincomb2: k1")
        a3, h3 = _as_buf(p2, n * pb, "lincomb2: p2")
        a4, h4 = _as_buf(k2, n * fb, "lincomb2: k2")
        po, h5 = _as_buf(out)
        pv, h6 = _as_buf(invalid)
        self._check(self.lib.ecb200_lincomb2(self.h, cid, n, a1, a2, a3, a4, po, pv, flags), "lincomb2")
        return bytes(out), bytes(invalid)

    def ecdsa_verify(self, curve, q: bytes, z: bytes, rs: bytes, out=None) -> bytes:
        """ok bytes for n x (Q = x||y, z = bits2field(prehash), r||s).  Buffers may be bytes or numpy uint8 arrays;
        page-locked arrays (e.g. views of torch pinned tensors) are DMA'd directly, pageable ones are staged.
        out: optional n-byte numpy array that receives the result (returned as is)."""
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(z, fb, "ecdsa_verify: z")
        ok = out if out is not None else bytearray(n)
        pq, a = _as_buf(q, n * 2 * fb, "ecdsa_verify: q")
        pz, b = _as_buf(z, n * fb, "ecdsa_verify: z")
        pr, c = _as_buf(rs, n * 2 * fb, "ecdsa_verify: rs")
        po, d = _as_buf(ok, n, "ecdsa_verify: out")
        self._check(self.lib.ecb200_ecdsa_verify(self.h, cid, n, pq, pz, pr, po), "ecdsa_verify")
        return ok if out is not None else bytes(ok)

    def verify_prehash_batch(self, curve, keys: Sequence[Tuple[int, int]], prehashes: Sequence[bytes],
                             sigs: Sequence[Tuple[int, int]]) -> List[bool]:
        """Vec<Result<(), Error>> of VerifyingKey::verify_prehash over slices (True = Ok(()))."""
        cid, fb = curve_id(curve), field_bytes(curve)
        if not (len(keys) == len(prehashes) == len(sigs)):
            raise ValueError("verify_prehash_batch: keys, prehashes and sigs must have the same length")
        lim = 1 << (8 * fb)
        idx, q, z, rs = [], bytearray(), bytearray(), bytearray()
        res = [False] * len(keys)
        for i, ((x, y), h, (r, s)) in enumerate(zip(keys, prehashes, sigs)):
            try:
                zb = bits2field(cid, h)
            except ValueError:
                continue
            if not (0 <= x < lim and 0 <= y < lim and 0 <= r < lim and 0 <= s < lim):
                continue
            idx.append(i)
            q += x.to_bytes(fb, "big") + y.to_bytes(fb, "big")
            z += zb
            rs += r.to_bytes(fb, "big") + s.to_bytes(fb, "big")
        ok = self.ecdsa_verify(cid, bytes(q), bytes(z), bytes(rs))
        for j, i in enumerate(idx):
            res[i] = bool(ok[j])
        return res

    def field_op(self, curve, which: int, op: int, a: bytes, b: Optional[bytes] = None) -> Tuple[bytes, bytes]:
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(a, fb, "field_op: a")
        out = bytearray(n * fb)
        ok = bytearray(n)
        pa, k1 = _as_buf(a, n * fb, "field_op: a")
        pb, k2 = _as_buf(b, n * fb, "field_op: b")
        po, k3 = _as_buf(out)
        pk, k4 = _as_buf(ok)
        self._check(self.lib.ecb200_field_op(self.h, cid, which, op, n, pa, pb, po, pk), "field_op")
        return bytes(out), bytes(ok)

    # ------------------------------------------------------------------ SURVEY §8f rows (host buffers)
    def decode_points(self, curve, enc: bytes, stride: int, mode: int = DECODE_SEC1) -> Tuple[bytes, bytes]:
        """from_encoded_point / decompress / decompact over n fixed-stride slots -> (x||y bytes, status bytes 1/2/0)."""
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(enc, stride, "decode_points: enc")
        xy, st = bytearray(n * 2 * fb), bytearray(n)
        pe, a = _as_buf(enc, n * stride, "decode_points: enc")
        px, b = _as_buf(xy)
        ps, c = _as_buf(st)
        self._check(self.lib.ecb200_decode_points(self.h, cid, n, pe, stride, mode, px, ps), "decode_points")
        return bytes(xy), bytes(st)

    def ecdsa_verify_sec1(self, curve, keys: bytes, key_stride: int, z: bytes, rs: bytes) -> bytes:
        """verify_prehash with SEC1-encoded keys (VerifyingKey::from_sec1_bytes), keys decoded on the device."""
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(z, fb, "ecdsa_verify_sec1: z")
        ok = bytearray(n)
        pk, a = _as_buf(keys, n * key_stride, "ecdsa_verify_sec1: keys")
        pz, b = _as_buf(z, n * fb, "ecdsa_verify_sec1: z")
        pr, c = _as_buf(rs, n * 2 * fb, "ecdsa_verify_sec1: rs")
        po, d = _as_buf(ok)
        self._check(self.lib.ecb200_ecdsa_verify_sec1(self.h, cid, n, pk, key_stride, pz, pr, po), "ecdsa_verify_sec1")
        return bytes(ok)

    def ecdsa_recover(self, curve, z: bytes, rs: bytes, recid: bytes, flags: int = 0) -> Tuple[bytes, bytes]:
        """VerifyingKey::recover_from_prehash over a batch -> (SEC1 key slots, ok bytes)."""
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _nbytes(recid)
        keys, ok = bytearray(n * slot_bytes(cid, flags)), bytearray(n)
        pz, a = _as_buf(z, n * fb, "ecdsa_recover: z")
        pr, b = _as_buf(rs, n * 2 * fb, "ecdsa_recover: rs")
        pi, c = _as_buf(recid, n, "ecdsa_recover: recid")
        pk, d = _as_buf(keys)
        po, e = _as_buf(ok)
        self._check(self.lib.ecb200_ecdsa_recover(self.h, cid, n, pz, pr, pi, pk, po, flags), "ecdsa_recover")
        return bytes(keys), bytes(ok)

    def schnorr_verify(self, pk: bytes, e: bytes, sig: bytes) -> bytes:
        """BIP340 verification after hashing: pk n x 32, e n x 32 challenge digests, sig n x 64."""
        n = _rows(pk, 32, "schnorr_verify: pk")
        ok = bytearray(n)
        pp, a = _as_buf(pk, n * 32, "schnorr_verify: pk")
        pe, b = _as_buf(e, n * 32, "schnorr_verify: e")
        ps, c = _as_buf(sig, n * 64, "schnorr_verify: sig")
        po, d = _as_buf(ok)
        self._check(self.lib.ecb200_schnorr_verify(self.h, n, pp, pe, ps, po), "schnorr_verify")
        return bytes(ok)

    def sm2dsa_verify(self, q: bytes, e: bytes, rs: bytes) -> bytes:
        n = _rows(e, 32, "sm2dsa_verify: e")
        ok = bytearray(n)
        pq, a = _as_buf(q, n * 64, "sm2dsa_verify: q")
        pe, b = _as_buf(e, n * 32, "sm2dsa_verify: e")
        pr, c = _as_buf(rs, n * 64, "sm2dsa_verify: rs")
        po, d = _as_buf(ok)
        self._check(self.lib.ecb200_sm2dsa_verify(self.h, n, pq, pe, pr, po), "sm2dsa_verify")
        return bytes(ok)

    def ecdsa_sign(self, curve, d: bytes, k: bytes, z: bytes) -> Tuple[bytes, bytes, bytes]:
        """try_sign_prehashed with caller-supplied nonces -> (r||s, recovery ids, ok)."""
        cid, fb = curve_id(curve), field_bytes(curve)
        n = _rows(z, fb, "ecdsa_sign: z")
        rs, rid, ok = bytearray(n * 2 * fb), bytearray(n), bytearray(n)
        pd, a = _as_buf(d, n * fb, "ecdsa_sign: d")
        pk, b = _as_buf(k, n * fb, "ecdsa_sign: k")
        pz, c = _as_buf(z, n * fb, "ecdsa_sign: z")
        pr, e = _as_buf(rs)
        pi, f = _as_buf(rid)
        po, g = _as_buf(ok)
        self._check(self.lib.ecb200_ecdsa_sign(self.h, cid, n, pd, pk, pz, pr, pi, po), "ecdsa_sign")
        return bytes(rs), bytes(rid), bytes(ok)

    # ------------------------------------------------------------------ device-pointer API (torch uint8 CUDA tensors)
    @staticmethod
    def _ptr(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def mul_gen_dev(self, curve, n, d_k, d_out, flags=0, stream=None):
        self._check(self.lib.ecb200_mul_gen_dev(self.h, curve_id(curve), n, self._ptr(d_k), self._ptr(d_out), flags,
                                                ctypes.c_void_p(stream) if stream else None), "mul_gen_dev")

    def mul_var_dev(self, curve, n, d_pts, d_inf, d_k, d_out, d_invalid=None, flags=0, stream=None):
        self._check(self.lib.ecb200_mul_var_dev(self.h, curve_id(curve), n, self._ptr(d_pts), self._ptr(d_inf), self._ptr(d_k),
                                                self._ptr(d_out), self._ptr(d_invalid), flags,
                                                ctypes.c_void_p(stream) if stream else None), "mul_var_dev")

    def batch_normalize_dev(self, curve, n, d_xyz, d_xy, d_inf=None, stream=None):
        self._check(self.lib.ecb200_batch_normalize_dev(self.h, curve_id(curve), n, self._ptr(d_xyz), self._ptr(d_xy),
                                                        self._ptr(d_inf), ctypes.c_void_p(stream) if stream else None),
                    "batch_normalize_dev")

    def ecdsa_verify_dev(self, curve, n, d_q, d_z, d_rs, d_ok, stream=None):
        self._check(self.lib.ecb200_ecdsa_verify_dev(self.h, curve_id(curve), n, self._ptr(d_q), self._ptr(d_z), self._ptr(d_rs),
                                                     self._ptr(d_ok), ctypes.c_void_p(stream) if stream else None),
                    "ecdsa_verify_dev")


    def lincomb2_dev(self, curve, n, d_p1, d_k1, d_p2, d_k2, d_out, d_invalid=None, flags=0, stream=None):
        self._check(self.lib.ecb200_lincomb2_dev(self.h, curve_id(curve), n, self._ptr(d_p1), self._ptr(d_k1), self._ptr(d_p2), self._ptr(d_k2),
                                                 self._ptr(d_out), self._ptr(d_invalid), flags, ctypes.c_void_p(stream) if stream else None),
                    "lincomb2_dev")

    def ecdsa_verify_sec1_dev(self, curve, n, d_keys, key_stride, d_z, d_rs, d_ok, stream=None):
        self._check(self.lib.ecb200_ecdsa_verify_sec1_dev(self.h, curve_id(curve), n, self._ptr(d_keys), key_stride, self._ptr(d_z),
                                                          self._ptr(d_rs), self._ptr(d_ok), ctypes.c_void_p(stream) if stream else None),
                    "ecdsa_verify_sec1_dev")

    def ecdsa_recover_dev(self, curve, n, d_z, d_rs, d_recid, d_keys, d_ok, flags=0, stream=None):
        self._check(self.lib.ecb200_ecdsa_recover_dev(self.h, curve_id(curve), n, self._ptr(d_z), self._ptr(d_rs), self._ptr(d_recid),
                                                      self._ptr(d_keys), self._ptr(d_ok), flags, ctypes.c_void_p(stream) if stream else None),
                    "ecdsa_recover_dev")

    def schnorr_verify_dev(self, n, d_pk, d_e, d_sig, d_ok, stream=None):
        self._check(self.lib.ecb200_schnorr_verify_dev(self.h, n, self._ptr(d_pk), self._ptr(d_e), self._ptr(d_sig), self._ptr(d_ok),
                                                       ctypes.c_void_p(stream) if stream else None), "schnorr_verify_dev")

    def sm2dsa_verify_dev(self, n, d_q, d_e, d_rs, d_ok, stream=None):
        self._check(self.lib.ecb200_sm2dsa_verify_dev(self.h, n, self._ptr(d_q), self._ptr(d_e), self._ptr(d_rs), self._ptr(d_ok),
                                                      ctypes.c_void_p(stream) if stream else None), "sm2dsa_verify_dev")

    def ecdsa_sign_dev(self, curve, n, d_d, d_k, d_z, d_rs, d_recid, d_ok, stream=None):
        self._check(self.lib.ecb200_ecdsa_sign_dev(self.h, curve_id(curve), n, self._ptr(d_d), self._ptr(d_k), self._ptr(d_z), self._ptr(d_rs),
                                                   self._ptr(d_recid), self._ptr(d_ok), ctypes.c_void_p(stream) if stream else None),
                    "ecdsa_sign_dev")

    def decode_points_dev(self, curve, n, d_enc, stride, mode, d_xy, d_status, stream=None):
        self._check(self.lib.ecb200_decode_points_dev(self.h, curve_id(curve), n, self._ptr(d_enc), stride, mode, self._ptr(d_xy),
                                                      self._ptr(d_status), ctypes.c_void_p(stream) if stream else None), "decode_points_dev")


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous index shard [lo, hi) of rank `rank`: floor(i*n/g) boundaries."""
    return (rank * n) // world, ((rank + 1) * n) // world
