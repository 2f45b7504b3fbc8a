"""-m gpu tests of the round-2 boundary additions, through the C ABI: the per-row two-term linear combination
(ecb200_lincomb2), one context over several devices (ecb200_init_multi), rejected lincomb terms, host-buffer validation of
the Python layer, and every index of the secret-scalar window table."""
import ctypes
import random

import numpy as np
import pytest

from oracle import ecoracle as o

pytestmark = pytest.mark.gpu
CUR = ["k256", "p256", "p384", "sm2", "p192", "p224"]


@pytest.fixture(scope="module")
def eng():
    import ecb200
    e = ecb200.Engine(0)
    yield e
    e.close()


def be(vals, fb):
    if isinstance(vals, int):
        vals = [vals]
    return b"".join(int(v).to_bytes(fb, "big") for v in vals)


def _pt(c, P):
    return be(P, c.fb)


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("ct", [0, 1])
def test_lincomb2_batch_vs_oracle(eng, cname, ct):
    """LinearCombination::lincomb(&x, &k, &y, &l) per row (k256 mul.rs:313-323, primeorder projective.rs:415-420; the
    reference's own test is mul.rs:493-507: lincomb == x*k + y*l).  Edge rows: cancellation (x*k + (-x)*k), y = x, zero
    scalars, k = n - 1, generator terms."""
    import ecb200
    c = o.curve(cname)
    fb = c.fb
    rng = random.Random(71 + c.cid)
    n = 150 if fb <= 32 else 60
    rows = []
    for i in range(n):
        x, y = o.mul_gen(c, rng.randrange(1, c.n)), o.mul_gen(c, rng.randrange(1, c.n))
        k, l = rng.randrange(c.n), rng.randrange(c.n)
        if i == 0: y, l = o.pt_neg(c, x), k                  # sum = identity
        if i == 1: y = x                                    # x*k + x*l
        if i == 2: k = 0
        if i == 3: k, l = 0, 0
        if i == 4: k, l = c.n - 1, 1
        if i == 5: x, y = c.G, c.G
        if i == 6: y, l = x, (c.n - k) % c.n                # x*k + x*(n-k) = identity
        rows.append((x, k, y, l))
    p1 = b"".join(_pt(c, r[0]) for r in rows); k1 = be([r[1] for r in rows], fb)
    p2 = b"".join(_pt(c, r[2]) for r in rows); k2 = be([r[3] for r in rows], fb)
    exp = b"".join(o.slot_encode(c, o.pt_lincomb(c, [(r[0], r[1]), (r[2], r[3])])) for r in rows)
    out, inv = eng.lincomb2_batch(cname, p1, k1, p2, k2, ct)
    assert out == exp and not any(inv)
    st = len(exp) // n
    assert out[:st] == bytes(st) and out[6 * st:7 * st] == bytes(st)
    # projective inputs with random Z
    def proj(P):
        lam = rng.randrange(1, c.p)
        return be((P[0] * lam % c.p, P[1] * lam % c.p, lam), fb)
    q1 = b"".join(proj(r[0]) for r in rows); q2 = b"".join(proj(r[2]) for r in rows)
    out, inv = eng.lincomb2_batch(cname, q1, k1, q2, k2, ct | ecb200.FLAG_PROJ)
    assert out == exp
    # an invalid point in either position voids that row only
    bad1 = be((1, 1), fb) + p1[2 * fb:6 * fb]
    bad2 = p2[:2 * fb] + be((c.p, 0), fb) + p2[4 * fb:6 * fb]
    out, inv = eng.lincomb2_batch(cname, bad1, k1[:3 * fb], bad2, k2[:3 * fb], ct)
    assert list(inv) == [1, 1, 0] and out[:2 * st] == bytes(2 * st) and out[2 * st:] == exp[2 * st:3 * st]
    assert eng.lincomb2_batch(cname, b"", b"", b"", b"", ct) == (b"", b"")


def test_lincomb2_matches_verify_construction(eng):
    """u1*G + u2*Q through lincomb2 equals the point ECDSA verification accepts on: R.x mod n == r for signed rows."""
    import ecb200
    c = o.K256
    rng = np.random.default_rng(5)
    n = 2000
    def scal():
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] &= 0x7F; a[:, -1] |= 1
        return a
    d, k, z = scal(), scal(), rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    rs, rid, ok = eng.ecdsa_sign("k256", d, k, z)
    assert ok == b"\x01" * n
    pub = np.frombuffer(eng.mul_by_generator_batch("k256", d, ecb200.FLAG_UNCOMPRESSED), np.uint8).reshape(n, 65)[:, 1:].copy()
    rsa = np.frombuffer(rs, np.uint8).reshape(n, 64)
    r, s = np.ascontiguousarray(rsa[:, :32]), np.ascontiguousarray(rsa[:, 32:])
    w, _ = eng.field_op("k256", 1, 5, s)
    zr, _ = eng.field_op("k256", 1, 0, z, bytes(n * 32))
    u1, _ = eng.field_op("k256", 1, 2, zr, w)
    u2, _ = eng.field_op("k256", 1, 2, r, w)
    g = np.tile(np.frombuffer(be(c.G, 32), np.uint8), (n, 1))
    out, inv = eng.lincomb2_batch("k256", g, u1, pub, u2, ecb200.FLAG_UNCOMPRESSED)
    x = np.frombuffer(out, np.uint8).reshape(n, 65)[:, 1:33]
    xr, _ = eng.field_op("k256", 1, 0, np.ascontiguousarray(x), bytes(n * 32))
    assert xr == r.tobytes()


@pytest.mark.parametrize("cname", ["k256", "p256"])
def test_lincomb_rejects_invalid_term(eng, cname):
    import ecb200
    c = o.curve(cname)
    fb = c.fb
    good = [o.mul_gen(c, 5), o.mul_gen(c, 9)]
    ks = be([3, 4], fb)
    assert eng.lincomb(cname, b"".join(_pt(c, P) for P in good), ks) == o.slot_encode(c, o.mul_gen(c, 51))
    for flags in (0, ecb200.FLAG_CT):
        with pytest.raises(ecb200.Ecb200Error, match="not a valid point"):
            eng.lincomb(cname, _pt(c, good[0]) + be((1, 1), fb), ks, flags)
    assert eng.lib.ecb200_lincomb(eng.h, c.cid, 2, ctypes.c_char_p(_pt(c, good[0]) + be((c.p, 0), fb)), ctypes.c_char_p(ks),
                                  ctypes.create_string_buffer(200), 0, 0) == -4


def _multi(devs):
    import ecb200
    return ecb200.Engine(devices=devs)


def test_init_multi_shards_inside_one_call(eng):
    """ecb200_init_multi: one context, several shards; every host entry point must return exactly what the single-device
    context returns, for sizes that split unevenly (and for n smaller than the number of shards).  On a one-GPU box the
    shards are child contexts on the same device; with >= 2 GPUs a second engine spans two devices."""
    import ecb200
    import torch
    c = o.K256
    rng = np.random.default_rng(11)
    cases = [[0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        cases.append([0, 1])
    for devs in cases:
        m = _multi(devs)
        assert m.n_devices == len(devs)
        for n in (0, 1, 2, 1001, 70001):
            ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
            a = eng.mul_by_generator_batch("k256", ks, ecb200.FLAG_UNCOMPRESSED)
            assert m.mul_by_generator_batch("k256", ks, ecb200.FLAG_UNCOMPRESSED) == a
            if n == 0:
                continue
            pts = np.frombuffer(a, np.uint8).reshape(n, 65)[:, 1:].copy()
            k2 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
            inf = np.zeros(n, np.uint8); inf[::7] = 1
            for fl in (0, ecb200.FLAG_CT):
                assert m.mul_batch("k256", pts, k2, inf, fl) == eng.mul_batch("k256", pts, k2, inf, fl)
            z = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
            d = ks.copy(); d[:, 0] &= 0x7F; d[:, -1] |= 1
            kk = k2.copy(); kk[:, 0] &= 0x7F; kk[:, -1] |= 1
            sig = eng.ecdsa_sign("k256", d, kk, z)
            assert m.ecdsa_sign("k256", d, kk, z) == sig
            pub = np.frombuffer(eng.mul_by_generator_batch("k256", d, ecb200.FLAG_UNCOMPRESSED), np.uint8).reshape(n, 65)[:, 1:].copy()
            rs = np.frombuffer(sig[0], np.uint8).reshape(n, 64).copy()
            rs[::5, 40] ^= 1
            okm = m.ecdsa_verify("k256", pub, z, rs)
            assert okm == eng.ecdsa_verify("k256", pub, z, rs)
            assert n < 5 or 0 < sum(okm) < n
            assert m.ecdsa_recover("k256", z, rs, sig[1]) == eng.ecdsa_recover("k256", z, rs, sig[1])
            assert m.lincomb2_batch("k256", pts, k2, pub, kk) == eng.lincomb2_batch("k256", pts, k2, pub, kk)
            xyz = np.concatenate([pts, z], axis=1); xyz[:, 64] &= 0x7F
            assert m.batch_normalize("k256", xyz) == eng.batch_normalize("k256", xyz)
            assert m.field_op("k256", 0, 2, z, kk) == eng.field_op("k256", 0, 2, z, kk)
            comp = eng.mul_by_generator_batch("k256", d)
            assert m.decode_points("k256", comp, 33) == eng.decode_points("k256", comp, 33)
            assert m.ecdsa_verify_sec1("k256", comp, 33, z, rs) == okm
            # many-term lincomb: per-device partials, summed on the first device
            if n <= 1001:
                assert m.lincomb("k256", pts, k2) == eng.lincomb("k256", pts, k2)
                assert m.lincomb("k256", pts, k2, ecb200.FLAG_CT) == eng.lincomb("k256", pts, k2)
        assert m.launch_count > 0
        # device pointers belong to one device
        t = torch.zeros(64, dtype=torch.uint8, device="cuda:0")
        with pytest.raises(ecb200.Ecb200Error, match="single-device"):
            m.mul_gen_dev("k256", 1, t, t)
        m.close()
    # other curves through the sharded path
    m = _multi([0, 0])
    for cname in ("p256", "p384", "sm2"):
        cc = o.curve(cname)
        n = 333
        ks = rng.integers(0, 256, size=(n, cc.fb), dtype=np.uint8)
        assert m.mul_by_generator_batch(cname, ks) == eng.mul_by_generator_batch(cname, ks)
    q = np.frombuffer(eng.mul_by_generator_batch("sm2", ks, ecb200.FLAG_UNCOMPRESSED), np.uint8).reshape(n, 65)[:, 1:].copy()
    e = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    rs = rng.integers(0, 256, size=(n, 64), dtype=np.uint8); rs[:, 0] &= 0x7F; rs[:, 32] &= 0x7F
    assert m.sm2dsa_verify(q, e, rs) == eng.sm2dsa_verify(q, e, rs)
    pk = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    assert m.schnorr_verify(pk, e, rs) == eng.schnorr_verify(pk, e, rs)
    m.close()
    assert eng.lib.ecb200_init_multi(1, (ctypes.c_int * 1)(99), ctypes.byref(ctypes.c_void_p())) == -1
    assert eng.lib.ecb200_device_count(eng.h) == 1


def test_two_contexts_two_host_threads(eng):
    """Two single-device contexts driven concurrently from two host threads: results and per-context launch counts are
    independent (launch counters are per context, the per-function attribute flags atomic)."""
    import threading
    import ecb200
    e2 = ecb200.Engine(0)
    rng = np.random.default_rng(3)
    ks = rng.integers(0, 256, size=(20000, 32), dtype=np.uint8)
    exp = eng.mul_by_generator_batch("k256", ks, ecb200.FLAG_CT)
    base = (eng.launch_count, e2.launch_count)
    res = {}
    def work(name, e, reps):
        for _ in range(reps):
            res[name] = e.mul_by_generator_batch("k256", ks, ecb200.FLAG_CT)
    ts = [threading.Thread(target=work, args=("a", eng, 6)), threading.Thread(target=work, args=("b", e2, 3))]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert res["a"] == exp and res["b"] == exp
    da, db = eng.launch_count - base[0], e2.launch_count - base[1]
    assert da == 2 * db and db > 0          # 6 vs 3 identical calls: counts are not shared between contexts
    e2.close()


def test_python_layer_validates_buffers(eng):
    """a short or strided buffer must raise before the C ABI is handed a raw pointer."""
    z = np.zeros((10, 32), np.uint8)
    q = np.zeros((10, 64), np.uint8)
    with pytest.raises(ValueError):
        eng.ecdsa_verify("k256", q[:9], z, q)                      # q shorter than n rows
    with pytest.raises(ValueError):
        eng.ecdsa_verify("k256", q, z, q[:, :63])                  # rs short and strided
    with pytest.raises(ValueError):
        eng.ecdsa_verify("k256", q[:, ::2], z[:, :16], q)          # strided views
    with pytest.raises(ValueError):
        eng.ecdsa_verify("k256", q.astype(np.uint16), z, q)        # wrong dtype
    with pytest.raises(ValueError):
        eng.mul_batch("k256", bytes(64 * 3), bytes(32 * 4))
    with pytest.raises(ValueError):
        eng.mul_batch("k256", bytes(64 * 4), bytes(32 * 4), inf=bytes(3))
    with pytest.raises(ValueError):
        eng.mul_by_generator_batch("p384", bytes(47))
    with pytest.raises(ValueError):
        eng.ecdsa_sign("k256", bytes(32), bytes(64), bytes(64))
    with pytest.raises(ValueError):
        eng.ecdsa_recover("k256", bytes(64), bytes(64), bytes(2))
    with pytest.raises(ValueError):
        eng.schnorr_verify(bytes(64), bytes(32), bytes(128))
    with pytest.raises(ValueError):
        eng.batch_normalize("k256", bytes(95))
    with pytest.raises(ValueError):
        eng.ecdsa_verify("k256", q, z, q, out=np.zeros(9, np.uint8))


@pytest.mark.parametrize("cname", CUR)
def test_every_window_table_index_secret_path(eng, cname):
    """ADVICE r1 (build_table): the constant-time multiplication with scalars whose signed radix-16 digits take every
    magnitude 0..8 and both signs in every window position - each entry of the 9-entry table is selected and used."""
    import ecb200
    c = o.curve(cname)
    fb = c.fb
    P = o.mul_gen(c, 0xABCDEF)
    ks = []
    for d in range(16):
        ks.append(int(("%x" % d) * (2 * fb), 16) % c.n)            # the same nibble in every window
        ks.append(d << (8 * fb - 8))
        ks.append(d)
    ks += [(c.n - 1) >> s for s in range(8)]
    pb = _pt(c, P) * len(ks)
    out, inv = eng.mul_batch(cname, pb, be(ks, fb), None, ecb200.FLAG_CT)
    assert out == o.batch_mul_var_affine(c, pb, None, be(ks, fb)) and not any(inv)
    out2, _ = eng.mul_batch(cname, pb, be(ks, fb), None, 0)
    assert out2 == out


def test_plain_c_multi_device_client(tmp_path):
    """A plain C program drives several shards (every GPU of the box, or two shards on device 0 when there is one GPU) from
    ONE process through ecb200_init_multi and compares every result with a single-device context."""
    import os
    import subprocess
    import ecb200
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cabi_multi_client")
    libdir = os.path.dirname(ecb200.LIB_PATH)
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cabi", "multi_client.c"),
                           "-o", exe, "-L", libdir, "-lecb200", "-Wl,-rpath," + libdir])
    shards = max(2, torch.cuda.device_count())
    out = subprocess.run([exe, str(shards)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "cabi multi client ok: %d shards" % shards in out.stdout


# ---------------------------------------------------------------------------------------------------------------------------
# per-key window tables of the verify path (rows grouped by public key inside a call; kernels.cuh "per-key window tables")

@pytest.fixture(scope="module")
def eng_rowpath():
    """an engine with the per-key tables switched off: the per-row path, for A/B comparisons"""
    import os
    import ecb200
    os.environ["ECB200_KEYTAB"] = "0"
    try:
        e = ecb200.Engine(0)
    finally:
        del os.environ["ECB200_KEYTAB"]
    yield e
    e.close()


def _wl():
    return __import__("importlib").import_module("rustcrypto-elliptic-curves_b200.workloads")


@pytest.mark.parametrize("cname", CUR)
def test_keytab_path_matches_per_row_path_and_construction(eng, eng_rowpath, cname):
    """Few keys, many rows: the table path must give exactly the mask of the per-row path and of the construction; one key is
    made invalid (off curve) for all of its rows, one group's rows all carry out-of-range signatures."""
    wl = _wl()
    c = o.curve(cname)
    fb = c.fb
    n, nk = (24000, 600) if fb <= 32 else (9000, 300)
    q, z, rs, exp = wl.make_verify_batch(wl.EngineBackend(eng, cname), cname, n, 0xB2100000 + c.cid, n_keys=nk)
    q, rs, exp = q.copy(), rs.copy(), exp.copy()
    bad_key = np.arange(n) % nk == 7
    q[bad_key, 2 * fb - 1] ^= 1                                   # the same broken bytes on every row of key 7
    exp[bad_key] = 0
    zero_r = np.arange(n) % nk == 11
    rs[zero_r, :fb] = 0                                            # r = 0 on every row of key 11
    exp[zero_r] = 0
    r0, t0 = eng.keytab_stats()
    got = np.frombuffer(eng.ecdsa_verify(cname, q, z, rs), np.uint8)
    r1, t1 = eng.keytab_stats()
    assert r1 - r0 == n and t1 - t0 == nk, "the per-key table path did not run"
    ref = np.frombuffer(eng_rowpath.ecdsa_verify(cname, q, z, rs), np.uint8)
    assert eng_rowpath.keytab_stats() == (0, 0)
    assert np.array_equal(got, ref) and np.array_equal(got, exp)
    assert 0 < int(got.sum()) < n
    # oracle on a sample (incl. rows of the broken groups)
    for i in list(range(0, n, n // 61)) + [7, 11, 7 + nk, 11 + nk, 5, 21, 37, 53, 69]:
        Q = (int.from_bytes(q[i, :fb].tobytes(), "big"), int.from_bytes(q[i, fb:].tobytes(), "big"))
        r, s = int.from_bytes(rs[i, :fb].tobytes(), "big"), int.from_bytes(rs[i, fb:].tobytes(), "big")
        assert bool(got[i]) == o.verify_prehashed(c, Q, z[i].tobytes(), r, s), i
    # permuting the rows permutes the mask (group numbering is arbitrary, results are not)
    perm = np.random.default_rng(3).permutation(n)
    gotp = np.frombuffer(eng.ecdsa_verify(cname, q[perm].copy(), z[perm].copy(), rs[perm].copy()), np.uint8)
    assert np.array_equal(gotp, got[perm])


def test_keytab_policy_and_chunked_host_calls(eng, eng_rowpath):
    """(a) all keys distinct: the per-row path runs (no tables); (b) a host call that spans several pipeline chunks carries the
    key groups from chunk to chunk (tables are built once per key per call); (c) below 4 rows per key the per-row path runs, from 4 on the (narrow) tables;
    (d) the device-pointer entry point takes the same path; (e) SEC1-encoded keys and SM2DSA go through it too."""
    import ecb200
    import torch
    wl = _wl()
    c = o.K256
    be_ = wl.EngineBackend(eng, "k256")
    n = 40000
    q, z, rs, exp = wl.make_verify_batch(be_, "k256", n, 0xB2100100, n_keys=n)          # (a)
    r0, t0 = eng.keytab_stats()
    assert eng.ecdsa_verify("k256", q, z, rs) == exp.tobytes()
    assert eng.keytab_stats() == (r0, t0)
    q, z, rs, exp = wl.make_verify_batch(be_, "k256", n, 0xB2100101, n_keys=n // 3)     # (c) 3 rows per key: below the policy's 4
    assert eng.ecdsa_verify("k256", q, z, rs) == exp.tobytes()
    assert eng.keytab_stats() == (r0, t0)
    q, z, rs, exp = wl.make_verify_batch(be_, "k256", n, 0xB2100103, n_keys=n // 5)     # 5 rows per key: narrow tables
    assert eng.ecdsa_verify("k256", q, z, rs) == exp.tobytes()
    assert eng.keytab_stats() == (r0 + n, t0 + n // 5)
    r0, t0 = eng.keytab_stats()
    n = 1500000                                                                            # (b) 227 328 + 1 136 640 + rest: three chunks
    nk = 3000
    q, z, rs, exp = wl.make_verify_batch(be_, "k256", n, 0xB2100102, n_keys=nk)
    got = eng.ecdsa_verify("k256", q, z, rs)
    r1, t1 = eng.keytab_stats()
    assert got == exp.tobytes()
    assert r1 - r0 == n and t1 - t0 == nk, (r1 - r0, t1 - t0)
    assert eng_rowpath.ecdsa_verify("k256", q, z, rs) == got
    dev = torch.device("cuda:0")                                                           # (d)
    m = 1 << 18
    dq, dz, drs = (torch.from_numpy(a[:m].copy()).to(dev) for a in (q, z, rs))
    dok = torch.empty(m, dtype=torch.uint8, device=dev)
    eng.ecdsa_verify_dev("k256", m, dq, dz, drs, dok)
    torch.cuda.synchronize()
    assert dok.cpu().numpy().tobytes() == exp[:m].tobytes()
    r2, t2 = eng.keytab_stats()
    assert r2 - r1 == m and t2 - t1 == nk
    # (e) compressed SEC1 keys: decoded on the device, then grouped
    d = wl.random_scalars(nk, 32, 0xB2100102)
    comp = np.frombuffer(eng.mul_by_generator_batch("k256", d), np.uint8).reshape(nk, 33)
    keys = np.ascontiguousarray(comp[np.arange(m) % nk])
    assert eng.ecdsa_verify_sec1("k256", keys, 33, z[:m].copy(), rs[:m].copy()) == exp[:m].tobytes()
    assert eng.keytab_stats()[0] - r2 == m
    # SM2DSA on tables vs per-row
    ns, nks = 20000, 100
    rng = np.random.default_rng(9)
    dk = wl.random_scalars(nks, 32, 77)
    pub = np.frombuffer(eng.mul_by_generator_batch("sm2", dk, ecb200.FLAG_UNCOMPRESSED), np.uint8).reshape(nks, 65)[:, 1:]
    qs = np.ascontiguousarray(pub[np.arange(ns) % nks])
    e = rng.integers(0, 256, size=(ns, 32), dtype=np.uint8)
    sg = rng.integers(0, 256, size=(ns, 64), dtype=np.uint8)
    sg[:, 0] &= 0x7F
    sg[:, 32] &= 0x7F
    assert eng.sm2dsa_verify(qs, e, sg) == eng_rowpath.sm2dsa_verify(qs, e, sg)


def test_keytab_wycheproof_and_crafted_rows_tiled(eng, golden):
    """The reference's own verification vectors (all Wycheproof rows) and the crafted exceptional-case rows (P + P, P + (-P),
    identity accumulators, x(R) >= n), tiled so that every key repeats and the table path runs: same verdicts as the oracle."""
    import hashlib
    from tests import crafted
    for cname in ("k256", "p256", "p384"):
        c = o.curve(cname)
        fb = c.fb
        blob = golden["wycheproof"][cname]
        hf = getattr(hashlib, blob["hash"])
        rows = []
        for wx, wy, msg, sig, flag in blob["rows"]:
            rsv = o.der_parse_strict(bytes.fromhex(sig), c)
            if rsv is None:
                continue
            Q = (int.from_bytes(bytes.fromhex(wx)[-fb:], "big"), int.from_bytes(bytes.fromhex(wy)[-fb:], "big"))
            zb = ecb_bits2field(cname, hf(bytes.fromhex(msg)).digest())
            if rsv[0] >= 1 << (8 * fb) or rsv[1] >= 1 << (8 * fb):
                continue
            rows.append((Q, zb, rsv[0], rsv[1]))
        rows += [(Q, zb, r, s) for Q, zb, r, s in crafted.exceptional_rows(c) + crafted.reduced_x_rows(c)]
        reps = 24
        q = b"".join(be(r[0], fb) for r in rows) * reps
        z = b"".join(r[1] for r in rows) * reps
        rs = b"".join(be((r[2], r[3]), fb) for r in rows) * reps
        exp = o.batch_verify(c, q[:len(rows) * 2 * fb], z[:len(rows) * fb], rs[:len(rows) * 2 * fb]) * reps
        r0, _ = eng.keytab_stats()
        got = eng.ecdsa_verify(cname, q, z, rs)
        assert eng.keytab_stats()[0] - r0 == len(rows) * reps, "the per-key table path did not run"
        assert got == exp and 0 < sum(got) < len(got)


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p224"])
def test_fixed_base_window_collisions_at_device_widths(eng, cname):
    """Rows whose accumulator u2*Q + (windows of u1*G below j) equals the table entry of window j or its negative
    (tests/crafted.py fixed_base_collision_rows), for the window width the device tables really have (20 bits on the 256-bit and
    smaller curves, 16 on P-384): the P + P and P + (-P) branches of the gathered mixed addition at every window position incl.
    the unsigned top window - on the per-row kernels (a small call) and on the per-key tables (the same rows tiled)."""
    from tests import crafted
    c = o.curve(cname)
    fb = c.fb
    gw = 16 if fb > 32 else 20
    rows = crafted.fixed_base_collision_rows(c, gw)
    nwin = (8 * fb + gw - 1) // gw
    assert len(rows) >= 2 * nwin - 2
    rows += [(Q, z, r % (c.n - 1) + 1, s) for Q, z, r, s in rows[::3]]           # same path, wrong r: must be rejected
    q = b"".join(be(r[0], fb) for r in rows)
    z = b"".join(r[1] for r in rows)
    rs = b"".join(be((r[2], r[3]), fb) for r in rows)
    exp = o.batch_verify(c, q, z, rs)
    assert sum(exp) >= 2 * nwin - 4 and sum(exp) < len(exp)
    r0, _ = eng.keytab_stats()
    assert eng.ecdsa_verify(cname, q, z, rs) == exp                             # < 4096 rows: per-row path
    assert eng.keytab_stats()[0] == r0
    reps = 4096 // len(rows) + 1
    got = eng.ecdsa_verify(cname, q * reps, z * reps, rs * reps)
    assert eng.keytab_stats()[0] - r0 == len(rows) * reps, "the per-key table path did not run"
    assert got == exp * reps


def ecb_bits2field(cname, digest):
    import ecb200
    return ecb200.bits2field(cname, digest)


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("shift", [1, 4, 8])
def test_device_pointers_of_any_alignment(eng, cname, shift):
    """The kernels move the ABI's byte strings as 128 / 64 / 32-bit words when the element address allows it and byte by byte
    otherwise (bigint.cuh load_be / store_be).  Every `_dev` entry point must give the same bytes for caller buffers at any
    offset from a cudaMalloc'd base: the same batch runs from 256-byte aligned tensors and from views shifted by 1, 4 and 8
    bytes (inputs AND outputs), through fixed-base, variable-base (CT and public), normalisation, signing and verification."""
    import torch
    import ecb200
    c = o.curve(cname)
    fb = c.fb
    n = 300
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(500 + c.cid)

    def scal(m):
        a = rng.integers(0, 256, size=(m, fb), dtype=np.uint8)
        a[:, 0] &= 0x3F
        a[:, -1] |= 1
        return a

    def put(a, sh):          # device copy of `a` starting `sh` bytes into a fresh allocation
        a = np.ascontiguousarray(a).reshape(-1)
        t = torch.zeros(a.size + 64, dtype=torch.uint8, device=dev)
        v = t[sh:sh + a.size]
        v.copy_(torch.from_numpy(a).to(dev))
        return v

    def room(size, sh):
        return torch.zeros(size + 64, dtype=torch.uint8, device=dev)[sh:sh + size]

    k, d, z, k2 = scal(n), scal(n), scal(n), scal(n)
    slot = 1 + 2 * fb
    res = {}
    for sh in (0, shift):
        r = {}
        # fixed base, CT, uncompressed slots
        out = room(n * slot, sh)
        eng.mul_gen_dev(cname, n, put(d, sh), out, ecb200.FLAG_CT | ecb200.FLAG_UNCOMPRESSED)
        torch.cuda.synchronize()
        r["gen"] = out.cpu().numpy().tobytes()
        pts = np.frombuffer(r["gen"], np.uint8).reshape(n, slot)[:, 1:]
        # variable base, public and secret scalars, compressed slots
        for ct in (0, ecb200.FLAG_CT):
            o2 = room(n * (1 + fb), sh)
            inv = room(n, sh)
            eng.mul_var_dev(cname, n, put(pts, sh), None, put(k2, sh), o2, inv, ct | ecb200.FLAG_COMPRESSED)
            torch.cuda.synchronize()
            r["var%d" % ct] = o2.cpu().numpy().tobytes() + inv.cpu().numpy().tobytes()
        # sign, then verify (valid rows and rows with a flipped message byte)
        rs, rid, ok = room(n * 2 * fb, sh), room(n, sh), room(n, sh)
        eng.ecdsa_sign_dev(cname, n, put(d, sh), put(k, sh), put(z, sh), rs, rid, ok)
        torch.cuda.synchronize()
        r["sign"] = rs.cpu().numpy().tobytes() + rid.cpu().numpy().tobytes() + ok.cpu().numpy().tobytes()
        zz = z.copy()
        zz[::3, fb - 1] ^= 1
        acc = room(n, sh)
        eng.ecdsa_verify_dev(cname, n, put(pts, sh), put(zz, sh), put(np.frombuffer(r["sign"][:n * 2 * fb], np.uint8), sh), acc)
        torch.cuda.synchronize()
        r["verify"] = acc.cpu().numpy().tobytes()
        res[sh] = r
    assert res[0] == res[shift]
    v = np.frombuffer(res[0]["verify"], np.uint8)
    assert v[1::3].all() and v[2::3].all() and not v[::3].any()
    # and the aligned run is right: first rows against the oracle
    for i in range(4):
        di = int.from_bytes(d[i].tobytes(), "big")
        assert res[0]["gen"][i * slot:(i + 1) * slot] == o.slot_encode(c, o.mul_gen(c, di), False)


@pytest.fixture(scope="module")
def eng_widths():
    """two engines with the per-key table width forced: narrow (4 bits) and wide (the curve's default: 6 bits on secp256k1,
    5 on P-256); abi.cu otherwise picks by rows per key (32 and more: wide)"""
    import os
    import ecb200
    es = []
    for v in ("0", "1"):
        os.environ["ECB200_KT_WIDE"] = v
        try:
            es.append(ecb200.Engine(0))
        finally:
            del os.environ["ECB200_KT_WIDE"]
    yield es
    for e in es:
        e.close()


@pytest.mark.parametrize("cname", ["k256", "p256", "sm2"])
def test_keytab_both_table_widths(eng, eng_rowpath, eng_widths, golden, cname):
    """The narrow and the wide per-key tables (different recodings, different kernels instantiations) give the mask of the per-row
    path on a synthetic batch with broken keys, and the oracle's verdicts on the reference's Wycheproof rows and the crafted
    exceptional-case rows tiled so that keys repeat; the automatic choice follows rows per key."""
    import hashlib
    from tests import crafted
    wl = _wl()
    c = o.curve(cname)
    fb = c.fb
    n, nk = 20000, 500
    q, z, rs, exp = wl.make_verify_batch(wl.EngineBackend(eng, cname), cname, n, 0xB2200000 + c.cid, n_keys=nk)
    q, exp = q.copy(), exp.copy()
    bad_key = np.arange(n) % nk == 3
    q[bad_key, fb - 1] ^= 1
    exp[bad_key] = 0
    ref = np.frombuffer(eng_rowpath.ecdsa_verify(cname, q, z, rs), np.uint8)
    assert np.array_equal(ref, exp)
    for e in eng_widths:
        r0, t0 = e.keytab_stats()
        got = np.frombuffer(e.ecdsa_verify(cname, q, z, rs), np.uint8)
        r1, t1 = e.keytab_stats()
        assert r1 - r0 == n and t1 - t0 == nk
        assert np.array_equal(got, ref)
    if cname == "sm2":
        return
    blob = golden["wycheproof"][cname]
    hf = getattr(hashlib, blob["hash"])
    rows = []
    for wx, wy, msg, sig, flag in blob["rows"]:
        rsv = o.der_parse_strict(bytes.fromhex(sig), c)
        if rsv is None or rsv[0] >= 1 << (8 * fb) or rsv[1] >= 1 << (8 * fb):
            continue
        Q = (int.from_bytes(bytes.fromhex(wx)[-fb:], "big"), int.from_bytes(bytes.fromhex(wy)[-fb:], "big"))
        rows.append((Q, ecb_bits2field(cname, hf(bytes.fromhex(msg)).digest()), rsv[0], rsv[1]))
    rows += [(Q, zb, r, s) for Q, zb, r, s in crafted.exceptional_rows(c) + crafted.reduced_x_rows(c)]
    reps = 24                                   # enough rows for the table policy (4096) on every curve
    qb = b"".join(be(r[0], fb) for r in rows) * reps
    zb = b"".join(r[1] for r in rows) * reps
    rsb = b"".join(be((r[2], r[3]), fb) for r in rows) * reps
    expw = o.batch_verify(c, qb[:len(rows) * 2 * fb], zb[:len(rows) * fb], rsb[:len(rows) * 2 * fb]) * reps
    for e in eng_widths:
        r0, _ = e.keytab_stats()
        assert e.ecdsa_verify(cname, qb, zb, rsb) == expw
        assert e.keytab_stats()[0] - r0 == len(rows) * reps, "the per-key table path did not run"


@pytest.mark.parametrize("gw", [7, 16, 18])
def test_fixed_base_table_widths(eng, gw):
    """The big fixed-base table of u1*G (jac.cuh add_fixed_base) at other window widths than the default 20 bits: 16 (the
    round-1 layout, word-aligned), 18 and 7 (windows straddling words, a short top window) - per-row path and per-key tables,
    the reference's Wycheproof rows plus a synthetic batch, on every 256-bit curve the table differs for."""
    import os
    import ecb200
    os.environ["ECB200_GW"] = str(gw)
    try:
        e = ecb200.Engine(0)
    finally:
        del os.environ["ECB200_GW"]
    try:
        wl = _wl()
        for cname in ("k256", "p256", "p384", "p224"):
            c = o.curve(cname)
            n, nk = 6000, 150
            q, z, rs, exp = wl.make_verify_batch(wl.EngineBackend(eng, cname), cname, n, 0xB2300000 + c.cid + gw, n_keys=nk)
            got = np.frombuffer(e.ecdsa_verify(cname, q, z, rs), np.uint8)
            assert np.array_equal(got, exp), (cname, "tables")
            m = 1500                               # below the table policy's minimum: per-row path
            got = np.frombuffer(e.ecdsa_verify(cname, q[:m].copy(), z[:m].copy(), rs[:m].copy()), np.uint8)
            assert np.array_equal(got, exp[:m]), (cname, "per-row")
            assert np.array_equal(got, np.frombuffer(eng.ecdsa_verify(cname, q[:m].copy(), z[:m].copy(), rs[:m].copy()), np.uint8))
    finally:
        e.close()
