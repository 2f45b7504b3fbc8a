"""Synthetic workloads of BASELINE.json's configs, generated with the engine itself.

Signatures are produced the way a signer does — R = k*G, r = x(R) mod n, s = k^-1 (z + r d) — with every
big-number step running as a batch on the GPU (mul_gen for k*G and the scalar-field hook for the mod-n
algebra), so 2^22 valid, distinct (Q, z, r, s) rows take about a second.  A deterministic 1/16 of the rows is
then corrupted; the expected accept mask follows from the construction and is what bench.py and the
large-size tests assert.  Nothing here touches oracle/.
"""
from __future__ import annotations

import numpy as np

from . import FLAG_UNCOMPRESSED, Engine, curve_id, field_bytes

ORDERS = {
    0: 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
    1: 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551,
    2: 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFC7634D81F4372DDF581A0DB248B0A77AECEC196ACCC52973,
    3: 0xFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFF7203DF6B21C6052B53BBF40939D54123,
    4: 0xFFFFFFFFFFFFFFFFFFFFFFFF99DEF836146BC9B1B4D22831,
    5: 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFF16A2E0B8F03E13DD29455C5C2A3D,
}
LOW_S = {0: True, 1: False, 2: False, 3: False, 4: False, 5: False}
N_KEYS = 1 << 16


def random_scalars(n: int, fb: int, seed: int) -> np.ndarray:
    """n x fb uniform bytes with the top bit cleared (< 2^(8fb-1) < n for every curve; never zero in practice)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, fb), dtype=np.uint8)
    a[:, 0] &= 0x7F
    a[:, fb - 1] |= 1
    return a


class EngineBackend:
    """Batch primitives the generator needs, on the GPU through the public host API."""

    def __init__(self, eng: Engine, curve):
        self.eng, self.cid, self.fb = eng, curve_id(curve), field_bytes(curve)

    def mul_gen_xy(self, k: np.ndarray) -> np.ndarray:
        n = k.shape[0]
        out = self.eng.mul_by_generator_batch(self.cid, k.tobytes(), FLAG_UNCOMPRESSED)
        return np.frombuffer(out, np.uint8).reshape(n, 1 + 2 * self.fb)[:, 1:]

    def fn(self, op: int, a: np.ndarray, b=None) -> np.ndarray:
        out, _ = self.eng.field_op(self.cid, 1, op, a.tobytes(), None if b is None else b.tobytes())
        return np.frombuffer(out, np.uint8).reshape(a.shape)


def make_verify_batch(backend, curve, n: int, seed: int, corrupt_every: int = 16, n_keys: int = N_KEYS):
    """Returns (q[n,2fb], z[n,fb], rs[n,2fb], expected[n]) as uint8 arrays.
    Row i uses key i mod n_keys (default 2^16 distinct keys; n_keys >= n makes every key distinct).  Rows with i % corrupt_every == 5 are corrupted (kind = (i // corrupt_every) % 5):
    0 flip a bit of r, 1 flip a bit of s, 2 flip a bit of z, 3 s -> n - s (high-s twin: rejected by k256 only),
    4 r -> r + n when that fits (out of range), else r = 0."""
    cid, fb = curve_id(curve), field_bytes(curve)
    order = ORDERS[cid]
    nk = min(n_keys, n)
    d = random_scalars(nk, fb, seed)
    qk = backend.mul_gen_xy(d)
    idx = np.arange(n) % nk
    q = np.ascontiguousarray(qk[idx])
    dd = np.ascontiguousarray(d[idx])
    k = random_scalars(n, fb, seed + 1)
    z = np.random.default_rng(seed + 2).integers(0, 256, size=(n, fb), dtype=np.uint8)
    R = backend.mul_gen_xy(k)
    zero = np.zeros((n, fb), np.uint8)
    r = backend.fn(0, np.ascontiguousarray(R[:, :fb]), zero)          # x mod n
    kinv = backend.fn(5, k)
    zr = backend.fn(0, z, zero)                                       # z mod n
    s = backend.fn(2, kinv, backend.fn(0, zr, backend.fn(2, r, dd)))  # k^-1 (z + r d)
    nbytes = np.frombuffer(order.to_bytes(fb, "big"), np.uint8)
    s_neg = backend.fn(4, s)                                          # n - s
    if LOW_S[cid]:
        half = np.frombuffer((order >> 1).to_bytes(fb, "big"), np.uint8)
        high = _gt(s, half)
        s = np.where(high[:, None], s_neg, s)
        s_neg = backend.fn(4, s)
    expected = np.ones(n, np.uint8)
    rows = np.nonzero(np.arange(n) % corrupt_every == 5)[0]
    kind = (rows // corrupt_every) % 5
    r = r.copy(); s = s.copy(); z = z.copy()
    rr = rows[kind == 0]; r[rr, fb - 1 - (rr % (fb - 1))] ^= (1 << (rr % 7)).astype(np.uint8)
    rr = rows[kind == 1]; s[rr, fb - 1 - (rr % (fb - 1))] ^= (1 << (rr % 7)).astype(np.uint8)
    rr = rows[kind == 2]; z[rr, rr % fb] ^= (1 << (rr % 7)).astype(np.uint8)
    expected[rows[kind <= 2]] = 0
    rr = rows[kind == 3]; s[rr] = s_neg[rr]
    expected[rr] = 0 if LOW_S[cid] else 1
    rr = rows[kind == 4]
    if len(rr):
        rn = _add_const(r[rr], nbytes)                                # r + n: out of range or wrapped (-> tiny, wrong)
        r[rr] = rn
        expected[rr] = 0
    rs = np.concatenate([r, s], axis=1)
    return q, np.ascontiguousarray(z), np.ascontiguousarray(rs), expected


def _gt(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """row-wise big-endian a > b (b a single row)."""
    diff = a != b[None, :]
    first = diff.argmax(axis=1)
    anyd = diff.any(axis=1)
    rows = np.arange(a.shape[0])
    return anyd & (a[rows, first] > b[first])


def _add_const(a: np.ndarray, c: np.ndarray) -> np.ndarray:
    """row-wise big-endian (a + c) mod 2^(8fb)."""
    out = np.empty_like(a)
    carry = np.zeros(a.shape[0], np.uint16)
    for j in range(a.shape[1] - 1, -1, -1):
        t = a[:, j].astype(np.uint16) + np.uint16(c[j]) + carry
        out[:, j] = (t & 0xFF).astype(np.uint8)
        carry = t >> 8
    return out


def make_mul_var_batch(backend, curve, n: int, seed: int, projective: bool = False):
    """(points, scalars): P_i = (seeded scalar)*G from the fixed-base kernel; scalars uniform.
    With projective=True points are X||Y||Z with a random non-unit Z (X = x*l, Y = y*l, Z = l),
    so batch normalisation is exercised (config 2)."""
    cid, fb = curve_id(curve), field_bytes(curve)
    base = random_scalars(n, fb, seed)
    xy = backend.mul_gen_xy(base)
    k = random_scalars(n, fb, seed + 1)
    if not projective:
        return np.ascontiguousarray(xy), k
    lam = random_scalars(n, fb, seed + 2)
    eng = backend.eng
    X, _ = eng.field_op(cid, 0, 2, np.ascontiguousarray(xy[:, :fb]).tobytes(), lam.tobytes())
    Y, _ = eng.field_op(cid, 0, 2, np.ascontiguousarray(xy[:, fb:]).tobytes(), lam.tobytes())
    lam_red, _ = eng.field_op(cid, 0, 0, lam.tobytes(), bytes(n * fb))
    xyz = np.concatenate([np.frombuffer(X, np.uint8).reshape(n, fb), np.frombuffer(Y, np.uint8).reshape(n, fb),
                          np.frombuffer(lam_red, np.uint8).reshape(n, fb)], axis=1)
    return np.ascontiguousarray(xyz), k
