"""Cross-checks against OpenSSL libcrypto (oracle/libcrypto_ref.py): an implementation that shares no code with the
oracle, the C++ port or the CUDA path.  CPU tier: oracle vs libcrypto (pins the oracle a second time; SM2 especially).
GPU tier (-m gpu): the CUDA path vs libcrypto at BASELINE config-1 size and on a synthetic verify batch."""
import random

import numpy as np
import pytest

from oracle import ecoracle as o

lc = pytest.importorskip("oracle.libcrypto_ref")
try:
    lc.lib()
    for _c in ("k256", "p256", "p384", "sm2", "p192", "p224"):
        lc.group(_c)
except Exception as e:  # pragma: no cover
    pytest.skip("libcrypto unusable: %s" % e, allow_module_level=True)

CUR = ["k256", "p256", "p384", "sm2", "p192", "p224"]


def be(v, n):
    return int(v).to_bytes(n, "big")


@pytest.mark.parametrize("cname", CUR)
def test_oracle_matches_libcrypto_mul(cname):
    c = o.curve(cname)
    rng = random.Random(41 + c.cid)
    ks = [0, 1, 2, c.n - 1, c.n - 2, c.n >> 1, c.n, c.n + 1, (1 << (8 * c.fb)) - 1] + [rng.randrange(c.n) for _ in range(40)]
    kb = b"".join(be(k % (1 << (8 * c.fb)), c.fb) for k in ks)
    for compress in (True, False):
        assert lc.mul_batch(cname, kb, None, compress, threads=2) == o.batch_mul_gen(c, kb, compress)
    pts = [o.mul_gen(c, rng.randrange(1, c.n)) for _ in ks]
    pb = b"".join(be(P[0], c.fb) + be(P[1], c.fb) for P in pts)
    assert lc.mul_batch(cname, kb, pb, None, threads=2) == o.batch_mul_var_affine(c, pb, None, kb)


@pytest.mark.parametrize("cname", CUR)
def test_oracle_matches_libcrypto_verify(cname):
    from tests import nextrows
    c = o.curve(cname)
    rng = random.Random(43 + c.cid)
    q, z, rs = bytearray(), bytearray(), bytearray()
    for i in range(30):
        d, k, zz, (r, s, _) = nextrows.make_sig(c, rng)
        Q = o.mul_gen(c, d)
        if i % 5 == 1:
            s = c.n - s
        if i % 5 == 2:
            r ^= 1
        if i % 7 == 3:
            Q = (Q[0], Q[1] ^ 1)
        if i % 11 == 4:
            s = c.n
        q += be(Q[0], c.fb) + be(Q[1], c.fb); z += zz; rs += be(r, c.fb) + be(s % (1 << 8 * c.fb), c.fb)
    got = lc.verify_batch(cname, bytes(q), bytes(z), bytes(rs), threads=2)
    assert got == o.batch_verify(c, bytes(q), bytes(z), bytes(rs))
    assert 0 < sum(got) < 30


@pytest.mark.gpu
def test_gpu_config1_vs_libcrypto():
    """BASELINE config 1 (k256 G*k, 2^16 scalars incl. the forced edge rows, 33-byte SEC1): 100 % vs libcrypto."""
    import ecb200
    c = o.K256
    n = 1 << 16
    rng = np.random.default_rng(0xB2000001)
    ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    edge = [0, 1, 2, c.n - 1, c.n - 2, c.n >> 1, (c.n >> 1) + 1, 1 << 128, (1 << 128) - 1]
    for i, v in enumerate(edge):
        ks[i] = np.frombuffer(be(v, 32), np.uint8)
    kb = ks.tobytes()
    eng = ecb200.Engine(0)
    got = eng.mul_by_generator_batch("k256", kb, ecb200.FLAG_CT)
    eng.close()
    assert got == lc.mul_batch("k256", kb)


@pytest.mark.gpu
@pytest.mark.parametrize("cname", CUR)
def test_gpu_verify_and_mul_vs_libcrypto(cname):
    import ecb200
    from tests import nextrows
    c = o.curve(cname)
    n = 4096 if c.fb == 32 else 1024
    rng = np.random.default_rng(0xB2000010 + c.cid)
    eng = ecb200.Engine(0)
    # sign on the device, corrupt a deterministic subset, verify on the device and with libcrypto
    def scal():
        a = rng.integers(0, 256, size=(n, c.fb), dtype=np.uint8)
        a[:, 0] &= 0x7F
        a[:, -1] |= 1
        return a
    d, k, z = scal(), scal(), rng.integers(0, 256, size=(n, c.fb), dtype=np.uint8)
    rs, rid, ok = eng.ecdsa_sign(cname, d.tobytes(), k.tobytes(), z.tobytes())
    assert ok == b"\x01" * n
    pub = eng.mul_by_generator_batch(cname, d.tobytes(), ecb200.FLAG_UNCOMPRESSED)
    q = np.frombuffer(pub, np.uint8).reshape(n, 1 + 2 * c.fb)[:, 1:].copy()
    rsa = np.frombuffer(rs, np.uint8).reshape(n, 2 * c.fb).copy()
    rsa[::5, 2 * c.fb - 1] ^= 1
    rsa[3::7, 5] ^= 0x40
    q[4::9, 2 * c.fb - 1] ^= 1
    got = eng.ecdsa_verify(cname, q.tobytes(), z.tobytes(), rsa.tobytes())
    assert got == lc.verify_batch(cname, q.tobytes(), z.tobytes(), rsa.tobytes())
    assert 0 < sum(got) < n
    # variable-base products
    out, inv = eng.mul_batch(cname, q.tobytes(), k.tobytes(), None, 0)
    assert out == lc.mul_batch(cname, k.tobytes(), q.tobytes())
    eng.close()
