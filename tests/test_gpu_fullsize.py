"""BASELINE-size parity of the scalar-multiplication configs (-m gpu, through the C ABI).

SURVEY.md section 8d, configs 2 and 5: k256 `P*k` + batch_normalize at 2^20 (projective X:Y:Z inputs with random Z and the
edge rows), p384 and sm2 `P*k` at 2^20.  Every row of the device output is compared with OpenSSL libcrypto
(oracle/libcrypto_ref.py mul_batch, an implementation that shares nothing with the engine, the oracle or the C++ port), a
2^16-row sample with oracle/ecoracle.py (the pinned restatement of the reference), and the constant-time kernels with the
variable-time ones on all rows.  The reference's own self-consistency tests for this path are k256/src/arithmetic/mul.rs:493-526
(lincomb / mul_by_generator vs Mul) and primeorder/src/dev.rs:7-157 (the MUL vectors, replayed in test_gpu_parity.py).

ECB200_TEST_LOG2 lowers the batch size for quick development runs (default 20 = the BASELINE size)."""
import os

import numpy as np
import pytest

from oracle import ecoracle as o
from tests import oracle_pool

pytestmark = pytest.mark.gpu

LOG2 = int(os.environ.get("ECB200_TEST_LOG2", "20"))
SAMPLE_LOG2 = int(os.environ.get("ECB200_TEST_SAMPLE_LOG2", "16"))

lc = pytest.importorskip("oracle.libcrypto_ref")


@pytest.fixture(scope="module")
def eng():
    import ecb200
    e = ecb200.Engine(0)
    yield e
    e.close()


def _be(v, fb):
    return np.frombuffer(int(v).to_bytes(fb, "big"), np.uint8)


def _edge_batch(eng, cname, n, seed):
    """(affine xy [n, 2fb], projective xyz [n, 3fb], scalars [n, fb], identity rows) of SURVEY 8d config 2:
    P_i = (seeded scalar) * G from the fixed-base kernel, projective form (x l, y l, l) with a seeded non-unit l,
    scalars uniform; forced edge rows: P = identity (Z = 0), k = 0, 1, n-1, P = G, P = -G, repeated P."""
    import ecb200
    wl = __import__("importlib").import_module("rustcrypto-elliptic-curves_b200.workloads")
    c = o.curve(cname)
    fb = c.fb
    xyz, k = wl.make_mul_var_batch(wl.EngineBackend(eng, cname), cname, n, seed, projective=True)
    xyz, k = xyz.copy(), k.copy()
    # recover the affine points the generator started from: libcrypto takes affine inputs
    xy, inf = eng.batch_normalize(cname, xyz)
    xy = np.frombuffer(xy, np.uint8).reshape(n, 2 * fb).copy()
    assert not any(inf)

    def set_proj(i, P, lam):
        xyz[i, :fb], xyz[i, fb:2 * fb], xyz[i, 2 * fb:] = _be(P[0] * lam % c.p, fb), _be(P[1] * lam % c.p, fb), _be(lam, fb)
        xy[i, :fb], xy[i, fb:] = _be(P[0], fb), _be(P[1], fb)

    ident = [0]
    xyz[0, :fb], xyz[0, fb:2 * fb], xyz[0, 2 * fb:] = _be(0, fb), _be(7, fb), _be(0, fb)      # P = identity (0 : y : 0)
    k[1] = _be(0, fb)
    k[2] = _be(1, fb)
    k[3] = _be(c.n - 1, fb)
    set_proj(4, c.G, 0x1234567)
    set_proj(5, o.pt_neg(c, c.G), c.p - 2)
    xyz[7], xy[7] = xyz[6], xy[6]                                                            # repeated P
    k[8] = _be(c.n >> 1, fb)
    k[9] = _be((1 << 128) - 1, fb)
    k[10] = _be(1 << 128, fb)
    return xy, np.ascontiguousarray(xyz), np.ascontiguousarray(k), ident


def _check_against_libcrypto(cname, got_slots, xy, k, ident, slot):
    n = k.shape[0]
    exp = np.frombuffer(lc.mul_batch(cname, k.tobytes(), xy.tobytes(), None), np.uint8).reshape(n, slot).copy()
    for i in ident:
        exp[i] = 0
    got = np.frombuffer(got_slots, np.uint8).reshape(n, slot)
    bad = np.nonzero((got != exp).any(axis=1))[0]
    assert bad.size == 0, "%d of %d rows differ from libcrypto, first %s" % (bad.size, n, bad[:5])


def _check_sample_against_oracle(cname, got_slots, xyz, k, slot, proj=True):
    c = o.curve(cname)
    n = k.shape[0]
    m = min(n, 1 << SAMPLE_LOG2)
    step = n // m
    idx = np.concatenate([np.arange(0, 16), np.arange(16, n, step)])[:m]     # the edge rows and an even spread
    exp = oracle_pool.mul_var(cname, c.fb, np.ascontiguousarray(xyz[idx]).tobytes(), np.ascontiguousarray(k[idx]).tobytes(), proj)
    got = np.frombuffer(got_slots, np.uint8).reshape(n, slot)[idx]
    assert got.tobytes() == exp


def test_config2_k256_mul_var_fullsize(eng):
    """BASELINE configs[1]: k256 variable-base `P*k` + batch_normalize, 2^20 (X:Y:Z, k) pairs, 33-byte SEC1 output."""
    import ecb200
    n = 1 << LOG2
    xy, xyz, k, ident = _edge_batch(eng, "k256", n, 0xB2000002)
    out_vt, inv = eng.mul_batch("k256", xyz, k, None, ecb200.FLAG_PROJ)
    assert not any(inv)
    out_ct, inv = eng.mul_batch("k256", xyz, k, None, ecb200.FLAG_PROJ | ecb200.FLAG_CT)
    assert not any(inv)
    assert out_ct == out_vt, "constant-time and variable-time kernels disagree"
    assert out_vt[:33] == bytes(33) and out_vt[33:66] == bytes(33)          # P = identity, k = 0
    _check_against_libcrypto("k256", out_vt, xy, k, ident, 33)
    _check_sample_against_oracle("k256", out_vt, xyz, k, 33)
    # the affine entry point (what a crate outside the fork can call) gives the same bytes
    out_aff, inv = eng.mul_batch("k256", xy[16:], k[16:], None, 0)
    assert out_aff == out_vt[16 * 33:]


@pytest.mark.parametrize("cname,seed", [("p384", 0xB2000005), ("sm2", 0xB2000006)])
def test_config5_primeorder_mul_var_fullsize(eng, cname, seed):
    """BASELINE configs[4]: p384 and sm2 variable-base scalar multiplication, 2^20 each, uncompressed SEC1."""
    import ecb200
    c = o.curve(cname)
    slot = 1 + 2 * c.fb
    n = 1 << LOG2
    xy, xyz, k, ident = _edge_batch(eng, cname, n, seed)
    out_vt, inv = eng.mul_batch(cname, xyz, k, None, ecb200.FLAG_PROJ)
    assert not any(inv)
    _check_against_libcrypto(cname, out_vt, xy, k, ident, slot)
    _check_sample_against_oracle(cname, out_vt, xyz, k, slot)
    out_ct, inv = eng.mul_batch(cname, xyz, k, None, ecb200.FLAG_PROJ | ecb200.FLAG_CT)
    assert not any(inv)
    assert out_ct == out_vt, "constant-time and variable-time kernels disagree"
    out_aff, inv = eng.mul_batch(cname, xy[16:], k[16:], None, 0)           # config 5 as bench.py times it: affine inputs
    assert out_aff == out_vt[16 * slot:]


def test_config1_k256_mul_gen_vs_oracle_sample(eng):
    """BASELINE configs[0] at its own size is checked 100 % against libcrypto in test_libcrypto_cross.py; here the same batch
    against the oracle on a 2^12 sample (the oracle needs ~3 ms per row), both output encodings."""
    import ecb200
    c = o.K256
    n = 1 << 16
    ks = np.random.default_rng(0xB2000001).integers(0, 256, size=(n, 32), dtype=np.uint8)
    for i, v in enumerate([0, 1, 2, c.n - 1, c.n - 2, c.n >> 1, (c.n >> 1) + 1, 1 << 128, (1 << 128) - 1]):
        ks[i] = _be(v, 32)
    got = np.frombuffer(eng.mul_by_generator_batch("k256", ks, ecb200.FLAG_CT), np.uint8).reshape(n, 33)
    idx = np.concatenate([np.arange(0, 16), np.arange(16, n, 16)])
    assert got[idx].tobytes() == oracle_pool.mul_gen("k256", 32, np.ascontiguousarray(ks[idx]).tobytes())
    got_u = np.frombuffer(eng.mul_by_generator_batch("k256", ks, ecb200.FLAG_UNCOMPRESSED), np.uint8).reshape(n, 65)
    assert np.array_equal(got_u[:, 1:33], got[:, 1:]) and np.array_equal(got_u[:, 0] == 4, got[:, 0] != 0)
