"""-m gpu tests of the split fixed-base path (kernels.cuh body_gen_half: two threads per scalar sum one half of the signed
5-bit windows each with Jacobian mixed additions, the halves are added with the complete formula inside the normalisation
kernel), through the C ABI.  Reference behaviour: ProjectivePoint::mul_by_generator (k256/src/arithmetic/mul.rs:424-439,
primeorder/src/projective.rs:422-431) followed by to_affine / to_encoded_point."""
import os
import random

import numpy as np
import pytest

from oracle import ecoracle as o
from oracle import libcrypto_ref as lc
from tests.test_emu_logic import split_gen_scalars

pytestmark = pytest.mark.gpu
CUR = ["k256", "p256", "p384", "sm2", "p192", "p224"]


@pytest.fixture(scope="module")
def eng():
    import ecb200
    e = ecb200.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_one_thread():
    """an engine with the split path switched off: one thread per scalar, complete mixed additions (round-1 kernel)"""
    import ecb200
    os.environ["ECB200_GEN2"] = "0"
    try:
        e = ecb200.Engine(0)
    finally:
        del os.environ["ECB200_GEN2"]
    yield e
    e.close()


def be(vals, fb):
    return b"".join(int(v).to_bytes(fb, "big") for v in vals)


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("ct", [0, 1])
def test_split_path_stress_scalars_vs_oracle(eng, cname, ct):
    """zero halves, single digits in every window position around the half boundary and the top window, carry chains of the
    signed recoding, unreduced inputs (k >= n is reduced once, as Scalar::reduce_bytes does)"""
    import ecb200
    c = o.curve(cname)
    rng = random.Random(0x5EED + c.cid)
    ks = split_gen_scalars(c, rng) + [rng.randrange(c.n) for _ in range(300)]
    # one row per (window, digit value): every table entry of every window is selected at least once
    W = 5
    nwin = 8 * c.fb // W + 1
    for w in range(nwin - 1):
        for v in range(1, 17):
            ks.append((v << (W * w)) % c.n)
            ks.append((c.n - (v << (W * w))) % c.n)
    kb = be([k % (1 << (8 * c.fb)) for k in ks], c.fb)
    flags = (ecb200.FLAG_CT if ct else 0)
    got = eng.mul_by_generator_batch(cname, kb, flags)
    exp = o.batch_mul_gen(c, kb)
    st = len(exp) // len(ks)
    bad = [i for i in range(len(ks)) if got[i * st:(i + 1) * st] != exp[i * st:(i + 1) * st]]
    assert not bad, (cname, ct, [hex(ks[i]) for i in bad[:4]])
    # uncompressed slots too
    got = eng.mul_by_generator_batch(cname, kb[:64 * c.fb], flags | ecb200.FLAG_UNCOMPRESSED)
    for i in range(64):
        P = o.mul_gen(c, ks[i] % c.n)
        slot = got[i * (1 + 2 * c.fb):(i + 1) * (1 + 2 * c.fb)]
        assert slot == (bytes(1 + 2 * c.fb) if P is None else b"\x04" + be(P, c.fb)), i
    assert eng.mul_by_generator_batch(cname, b"", flags) == b""


@pytest.mark.parametrize("cname", CUR)
def test_split_path_matches_one_thread_path(eng, eng_one_thread, cname):
    """same bytes from the two fixed-base kernels on a ragged batch (n not a multiple of the CTA size), CT and public"""
    import ecb200
    c = o.curve(cname)
    n = 20011 if c.fb <= 32 else 6007
    rng = np.random.default_rng(0xF1BA5E + c.cid)
    ks = rng.integers(0, 256, size=(n, c.fb), dtype=np.uint8)
    ks[:8] = 0
    ks[5, -1] = 1
    ks[9:12] = 0xFF                                                # >= n: reduced once
    kb = ks.tobytes()
    for flags in (0, ecb200.FLAG_CT, ecb200.FLAG_CT | ecb200.FLAG_UNCOMPRESSED):
        assert eng.mul_by_generator_batch(cname, kb, flags) == eng_one_thread.mul_by_generator_batch(cname, kb, flags), (cname, flags)


@pytest.mark.parametrize("cname", CUR)
def test_split_path_vs_libcrypto(eng, cname):
    import ecb200
    c = o.curve(cname)
    n = 1 << 14 if c.fb <= 32 else 1 << 12
    rng = np.random.default_rng(0xF1BA5F + c.cid)
    ks = rng.integers(0, 256, size=(n, c.fb), dtype=np.uint8)
    ks[:, 0] &= 0x7F
    kb = ks.tobytes()
    got = eng.mul_by_generator_batch(cname, kb, ecb200.FLAG_CT | ecb200.FLAG_COMPRESSED)
    assert got == lc.mul_batch(cname, kb, compress=True)


def test_signing_on_split_path(eng, eng_one_thread):
    """ecdsa_sign rides on the same fixed-base kernel: identical signatures from both kernels, and they verify"""
    for cname in ("k256", "p256", "p384"):
        c = o.curve(cname)
        n = 5000
        rng = np.random.default_rng(0x51C0 + c.cid)

        def scal():
            a = rng.integers(0, 256, size=(n, c.fb), dtype=np.uint8)
            a[:, 0] &= 0x7F
            a[:, -1] |= 1
            return a.tobytes()
        d, k, z = scal(), scal(), rng.integers(0, 256, size=(n, c.fb), dtype=np.uint8).tobytes()
        a = eng.ecdsa_sign(cname, d, k, z)
        b = eng_one_thread.ecdsa_sign(cname, d, k, z)
        assert a == b and a[2] == b"\x01" * n
        import ecb200
        pub = eng.mul_by_generator_batch(cname, d, ecb200.FLAG_UNCOMPRESSED)
        q = np.frombuffer(pub, np.uint8).reshape(n, 1 + 2 * c.fb)[:, 1:].tobytes()
        assert eng.ecdsa_verify(cname, q, z, a[0]) == b"\x01" * n


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("which", [0, 1])
def test_divsteps_inverse_on_device(eng, cname, which):
    """safegcd.cuh (the inversion of the Montgomery-trick kernels) against pow(a, m - 2, m) and against the Fermat chain on the
    device, base and scalar field of every curve: edge values, limb-boundary patterns, random values; 0 -> 0"""
    c = o.curve(cname)
    m = c.p if which == 0 else c.n
    fb = c.fb
    rng = random.Random(0xD1F5 + 2 * c.cid + which)
    vals = [0, 1, 2, 3, m - 1, m - 2, (m - 1) // 2, (m + 1) // 2]
    vals += [1 << k for k in range(0, 8 * fb - 1, 5)] + [(1 << k) - 1 for k in range(29, 8 * fb, 30)] + [m - (1 << k) for k in range(1, 8 * fb - 2, 13)]
    vals = [v % m for v in vals] + [rng.randrange(m) for _ in range(3000)]
    a = np.frombuffer(be(vals, fb), np.uint8).reshape(len(vals), fb)
    got, ok = eng.field_op(cname, which, 7, a, None)
    ref, _ = eng.field_op(cname, which, 5, a, None)
    assert bytes(ok) == b"\x01" * len(vals)
    assert bytes(got) == bytes(ref)
    assert bytes(got) == be([pow(v, m - 2, m) for v in vals], fb)
