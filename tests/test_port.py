"""Pins the C++ port (oracle/ecport.cpp, the timed CPU baseline) against the oracle: reference MUL
vectors, Wycheproof rows that reach arithmetic, random inputs.  CPU only."""
import ctypes
import hashlib
import random

import pytest

from oracle import ecoracle as o
from tests import port_lib

lib = port_lib.load()
CUR = ["k256", "p256", "p384", "sm2"]


def be(vals, fb):
    return b"".join(v.to_bytes(fb, "big") for v in vals)


@pytest.mark.parametrize("cname", CUR)
def test_port_mul_gen_and_var(cname, golden):
    c = o.curve(cname)
    fb = c.fb
    rng = random.Random(9 + c.cid)
    ks = [0, 1, 2, c.n - 1, c.n >> 1, (c.n >> 1) + 1, 1 << 128, c.n, c.n + 3] + [rng.randrange(c.n) for _ in range(40)]
    if cname in golden["group"]:
        ks += [int(k, 16) for k, _, _ in golden["group"][cname]["mul"]]
    kb = be([k % (1 << 8 * fb) for k in ks], fb)
    for comp in (0, 1):
        slot = 1 + (fb if comp else 2 * fb)
        out = (ctypes.c_uint8 * (len(ks) * slot))()
        assert lib.port_mul_gen(c.cid, len(ks), kb, out, comp) == 0
        assert bytes(out) == o.batch_mul_gen(c, kb, bool(comp))
    pts = [o.mul_gen(c, rng.randrange(1, c.n)) for _ in ks]
    pb = b"".join(be(P, fb) for P in pts)
    slot = 1 + 2 * fb
    out = (ctypes.c_uint8 * (len(ks) * slot))()
    assert lib.port_mul_var(c.cid, len(ks), pb, kb, out, 0) == 0
    assert bytes(out) == o.batch_mul_var_affine(c, pb, None, kb, False)


@pytest.mark.parametrize("cname", ["k256", "p256", "p384"])
def test_port_verify_wycheproof(cname, golden):
    c = o.curve(cname)
    blob = golden["wycheproof"][cname]
    hf = getattr(hashlib, blob["hash"])
    q, z, rs = bytearray(), bytearray(), bytearray()
    for wx, wy, msg, sig, flag in blob["rows"]:
        r_s = o.der_parse_strict(bytes.fromhex(sig), c)
        if r_s is None or r_s[0] >> (8 * c.fb) or r_s[1] >> (8 * c.fb):
            continue
        r, s = r_s
        q += bytes.fromhex(wx)[-c.fb:].rjust(c.fb, b"\0") + bytes.fromhex(wy)[-c.fb:].rjust(c.fb, b"\0")
        z += o.bits2field(c, hf(bytes.fromhex(msg)).digest())
        rs += be((r, s), c.fb)
        if c.low_s and 1 <= s < c.n and s > c.n >> 1:
            q += q[-2 * c.fb:]
            z += z[-c.fb:]
            rs += be((r, c.n - s), c.fb)
    n = len(z) // c.fb
    ok = (ctypes.c_uint8 * n)()
    assert lib.port_verify(c.cid, n, bytes(q), bytes(z), bytes(rs), ok) == 0
    assert bytes(ok) == o.batch_verify(c, bytes(q), bytes(z), bytes(rs))
    assert sum(ok) > 100


def test_port_verify_sm2_curve_ecdsa():
    c = o.SM2
    rng = random.Random(4)
    q, z, rs = bytearray(), bytearray(), bytearray()
    for i in range(12):
        d = rng.randrange(1, c.n)
        Q = o.mul_gen(c, d)
        zb = rng.randrange(1 << 256).to_bytes(32, "big")
        k = rng.randrange(1, c.n)
        r = o.mul_gen(c, k)[0] % c.n
        s = pow(k, -1, c.n) * (o.reduce_once(c, int.from_bytes(zb, "big")) + r * d) % c.n
        if i % 3 == 2:
            s ^= 4
        q += be(Q, 32); z += zb; rs += be((r, s), 32)
    ok = (ctypes.c_uint8 * 12)()
    lib.port_verify(c.cid, 12, bytes(q), bytes(z), bytes(rs), ok)
    assert bytes(ok) == o.batch_verify(c, bytes(q), bytes(z), bytes(rs))
    assert sum(ok) == 8
