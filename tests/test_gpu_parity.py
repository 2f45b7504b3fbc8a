"""Parity tests proper: the CUDA path, called through the C ABI (libecb200.so), against the oracle,
the committed golden vectors and size-independent properties.  Needs a B200: run with -m gpu."""
import hashlib
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


def H(s):
    return int(s, 16)


def be(vals, fb):
    return b"".join(v.to_bytes(fb, "big") for v in vals)


def ints(raw, fb):
    return [int.from_bytes(raw[i:i + fb], "big") for i in range(0, len(raw), fb)]


def edge_values(m):
    vals = [0, 1, 2, 3, m - 1, m - 2, m - 3, (m - 1) // 2, (m + 1) // 2, 2**32, 2**32 - 1, 2**32 + 977, 2**64 - 1,
            2**128, 2**128 - 1, 2**255 % m, (2**256 - 1) % m, 2**224 % m, (2**192 - 1) % m, 0xFFFFFFFF, m >> 32, m - 2**32]
    return sorted(set(v % m for v in vals))


# ------------------------------------------------------------------------------------------ fields
@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("which", [0, 1])
def test_field_ops_random_and_edges(eng, cname, which):
    """proptest analogue of k256 field.rs:792-872 / scalar.rs:1176-1281, p256 field.rs:787-803."""
    c = o.curve(cname)
    m = c.p if which == 0 else c.n
    rng = random.Random(100 + which * 10 + c.cid)
    ev = edge_values(m)
    A = [x for x in ev for _ in ev] + [rng.randrange(m) for _ in range(20000)]
    B = [y for _ in ev for y in ev] + [rng.randrange(m) for _ in range(20000)]
    a, b = be(A, c.fb), be(B, c.fb)
    for op, fn in ((0, lambda x, y: (x + y) % m), (1, lambda x, y: (x - y) % m), (2, lambda x, y: x * y % m)):
        out, ok = eng.field_op(cname, which, op, a, b)
        assert all(ok)
        assert ints(out, c.fb) == [fn(x, y) for x, y in zip(A, B)], (cname, which, op)
    out, _ = eng.field_op(cname, which, 3, a)
    assert ints(out, c.fb) == [x * x % m for x in A]
    out, _ = eng.field_op(cname, which, 4, a)
    assert ints(out, c.fb) == [(-x) % m for x in A]
    A2 = ev + [rng.randrange(m) for _ in range(500)]
    out, _ = eng.field_op(cname, which, 5, be(A2, c.fb))
    assert ints(out, c.fb) == [pow(x, m - 2, m) for x in A2]
    _, ok = eng.field_op(cname, which, 0, be([m, 2**(8 * c.fb) - 1], c.fb), be([0, 0], c.fb))
    assert list(ok) == [0, 0]


def test_field_golden_vectors(eng, golden):
    """DBL_TEST_VECTORS (k256 field.rs:663-689, p256 field.rs:724-749) and the risc0 8x32 KATs
    (field_8x32_risc0.rs:225-303) on the device."""
    for cname in ("k256", "p256"):
        c = o.curve(cname)
        dbl = [H(h) for h in golden["field"][cname]["dbl"]]
        out, ok = eng.field_op(cname, 0, 0, be(dbl[:-1], c.fb), be(dbl[:-1], c.fb))
        assert ints(out, c.fb) == dbl[1:] and all(ok)
        out, _ = eng.field_op(cname, 0, 2, be(dbl[:-1], c.fb), be([2] * (len(dbl) - 1), c.fb))
        assert ints(out, c.fb) == dbl[1:]
    k = golden["field"]["k256_risc0_8x32"]
    a, b = be([H(k["a"])], 32), be([H(k["b"])], 32)
    assert eng.field_op("k256", 0, 0, a, b)[0].hex() == k["add"]
    assert eng.field_op("k256", 0, 2, a, b)[0].hex() == k["mul"]
    assert eng.field_op("k256", 0, 3, a)[0].hex() == k["square"]
    assert eng.field_op("k256", 0, 4, a)[0].hex() == k["negate"]
    na = eng.field_op("k256", 0, 4, a)[0]
    nb = eng.field_op("k256", 0, 4, b)[0]
    assert eng.field_op("k256", 0, 0, na, nb)[0].hex() == k["add_negated"]
    zero = bytes(32)
    one = be([1], 32)
    assert eng.field_op("k256", 0, 2, a, zero)[0] == zero      # mul_zero
    assert eng.field_op("k256", 0, 2, a, one)[0] == a          # mul_one


@pytest.mark.parametrize("cname", CUR)
def test_sqrt(eng, cname):
    c = o.curve(cname)
    rng = random.Random(5)
    xs = [rng.randrange(c.p) for _ in range(300)] + list(range(1, 65))     # sqrt(1..64): primeorder field.rs:563-660
    sq = [x * x % c.p for x in xs]
    out, ok = eng.field_op(cname, 0, 6, be(sq, c.fb))
    assert all(ok)
    assert all(g * g % c.p == s for g, s in zip(ints(out, c.fb), sq))
    nonres = [v for v in range(2, 60) if pow(v, (c.p - 1) // 2, c.p) != 1][:8]
    _, ok = eng.field_op(cname, 0, 6, be(nonres, c.fb))
    assert not any(ok)


# ------------------------------------------------------------------------------------------ group KATs
@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p192", "p224"])
@pytest.mark.parametrize("flags", [0, 1])
def test_group_golden_vectors(eng, golden, cname, flags):
    """ADD_TEST_VECTORS ((i+1)*G) and MUL_TEST_VECTORS through mul_gen and mul_var (k256
    projective.rs:859-967; primeorder/src/dev.rs:7-157)."""
    c = o.curve(cname)
    fb = c.fb
    g = golden["group"][cname]
    ks = [i + 1 for i in range(len(g["add"]))] + [H(k) for k, _, _ in g["mul"]]
    exp_pts = [(H(x), H(y)) for x, y in g["add"]] + [(H(x), H(y)) for _, x, y in g["mul"]]
    exp = b"".join(o.slot_encode(c, P, False) for P in exp_pts)
    kb = be(ks, fb)
    assert eng.mul_by_generator_batch(cname, kb, flags | 4) == exp
    gb = be([c.gx, c.gy], fb) * len(ks)
    out, inv = eng.mul_batch(cname, gb, kb, None, flags | 4)
    assert out == exp and not any(inv)


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("ct", [0, 1])
def test_mul_gen_random_vs_oracle(eng, cname, ct):
    c = o.curve(cname)
    rng = random.Random(31 + c.cid)
    n = 300 if c.fb == 32 else 120
    ks = [0, 1, 2, c.n - 1, c.n - 2, c.n >> 1, (c.n >> 1) + 1, 1 << 128, (1 << 128) - 1, c.n, c.n + 5] + \
         [rng.randrange(c.n) for _ in range(n)]
    kb = be([k % (1 << 8 * c.fb) for k in ks], c.fb)
    for fl in (0, 2, 4):
        got = eng.mul_by_generator_batch(cname, kb, fl | ct)
        comp = True if fl == 2 else False if fl == 4 else None
        assert got == o.batch_mul_gen(c, kb, comp)


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("ct", [0, 1])
def test_mul_var_random_vs_oracle(eng, cname, ct):
    """k*P with affine and projective (random Z) inputs, edge rows: P = identity, k = 0, 1, n-1, P = +-G."""
    c = o.curve(cname)
    fb = c.fb
    rng = random.Random(41 + c.cid)
    n = 200 if fb == 32 else 80
    ks = [0, 1, 2, c.n - 1, c.n >> 1, 1 << 128, 7, 9] + [rng.randrange(c.n) for _ in range(n)]
    pts = [o.mul_gen(c, rng.randrange(1, c.n)) for _ in ks]
    pts[5] = c.G
    pts[6] = o.pt_neg(c, c.G)
    pts[7] = pts[8]                      # repeated P
    inf = bytearray(len(ks))
    inf[4] = 1
    pb = b"".join(be(P, fb) for P in pts)
    kb = be(ks, fb)
    out, invalid = eng.mul_batch(cname, pb, kb, bytes(inf), ct)
    assert out == o.batch_mul_var_affine(c, pb, bytes(inf), kb)
    assert not any(invalid)
    xyz = bytearray()
    for i, P in enumerate(pts):
        lam = rng.randrange(1, c.p)
        xyz += be((0, 5, 0), fb) if i == 3 else be((P[0] * lam % c.p, P[1] * lam % c.p, lam), fb)
    out, _ = eng.mul_batch(cname, bytes(xyz), kb, None, ct | 8)
    assert out == o.batch_mul_var_proj(c, bytes(xyz), kb)
    # invalid points: off-curve, coordinate >= p
    bad = be((1, 1), fb) + be((c.p, 0), fb) + be(c.G, fb)
    out, invalid = eng.mul_batch(cname, bad, be([5, 5, 5], fb), None, ct)
    st = len(out) // 3
    assert list(invalid) == [1, 1, 0]
    assert out[:2 * st] == bytes(2 * st) and out[2 * st:] == o.slot_encode(c, o.mul_gen(c, 5))


@pytest.mark.parametrize("cname", CUR)
def test_batch_normalize(eng, cname):
    """k256 projective.rs:773-834 (incl. IDENTITY slots); ragged size, n = 1."""
    c = o.curve(cname)
    fb = c.fb
    rng = random.Random(51)
    for n in (1, 37, 1000):
        pts = []
        for i in range(n):
            if n > 1 and i % 13 == 0:
                pts.append((rng.randrange(c.p), rng.randrange(1, c.p), 0))
            else:
                P = o.mul_gen(c, rng.randrange(1, 1 << 40))
                lam = rng.randrange(1, c.p)
                pts.append((P[0] * lam % c.p, P[1] * lam % c.p, lam))
        xyz = b"".join(be(P, fb) for P in pts)
        xy, inf = eng.batch_normalize(cname, xyz)
        exp = o.batch_normalize(c, pts)
        for i, P in enumerate(exp):
            sl = xy[i * 2 * fb:(i + 1) * 2 * fb]
            if P is None:
                assert inf[i] == 1 and sl == bytes(2 * fb)
            else:
                assert inf[i] == 0 and sl == be(P, fb)
    assert eng.batch_normalize(cname, b"") == (b"", b"")


@pytest.mark.parametrize("cname", ["k256", "p256", "p384"])
def test_batch_normalize_shared_inversion(eng, cname):
    """Several rows per thread and a ragged last CTA: the CTA-wide inversion (kernels_impl.cuh BlockInv: warp prefix /
    suffix products, one chain per CTA) must give every row its own 1/Z - including rows next to Z = 0 slots, which
    enter the shared product as 1 (BatchInvert semantics, k256 projective.rs:350-379)."""
    c = o.curve(cname)
    fb = c.fb
    rng = random.Random(52)
    base = [o.mul_gen(c, rng.randrange(1, c.n)) for _ in range(8)]
    n = 148 * 128 * 2 + 77                      # 3 rows per thread, last CTA partly idle
    pts = []
    for i in range(n):
        P = base[i % 8]
        lam = rng.randrange(1, c.p)
        if i % 129 == 5 or i in (0, n - 1):     # identity slots in every lane position over the batch
            pts.append((P[0] * lam % c.p, P[1] * lam % c.p, 0))
        else:
            pts.append((P[0] * lam % c.p, P[1] * lam % c.p, lam))
    xy, inf = eng.batch_normalize(cname, b"".join(be(P, fb) for P in pts))
    for i, P in enumerate(pts):
        sl = xy[i * 2 * fb:(i + 1) * 2 * fb]
        if P[2] == 0:
            assert inf[i] == 1 and sl == bytes(2 * fb), i
        else:
            assert inf[i] == 0 and sl == be(base[i % 8], fb), i


@pytest.mark.parametrize("cname", CUR)
def test_lincomb(eng, cname):
    """lincomb == sum k_i P_i (k256 mul.rs:493-526), incl. cancellation to the identity."""
    c = o.curve(cname)
    fb = c.fb
    rng = random.Random(61)
    for n in (1, 2, 3, 300):
        pts = [o.mul_gen(c, rng.randrange(1, 1 << 64)) for _ in range(n)]
        ks = [rng.randrange(c.n) for _ in range(n)]
        exp = o.pt_lincomb(c, list(zip(pts, ks)))
        got = eng.lincomb(cname, b"".join(be(P, fb) for P in pts), be(ks, fb))
        assert got == o.slot_encode(c, exp)
    P = o.mul_gen(c, 12345)
    got = eng.lincomb(cname, be(P, fb) + be(o.pt_neg(c, P), fb), be([77, 77], fb))
    assert got == o.slot_encode(c, None)
    # projective partial output can be fed back (multi-GPU reduction path)
    part = eng.lincomb(cname, be(P, fb), be([5], fb), 0, True)
    again = eng.lincomb(cname, part + part, be([1, 2], fb), 8)
    assert again == o.slot_encode(c, o.pt_mul(c, 15, P))


# ------------------------------------------------------------------------------------------ ECDSA
@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p192", "p224"])
def test_ecdsa_kats(eng, golden, cname):
    """new_verification_test!: verify OK; flip bit 0 of s[0] => Err (p256/src/ecdsa.rs:184-192)."""
    c = o.curve(cname)
    keys, hs, sigs, exp = [], [], [], []
    for v in golden["ecdsa"][cname]["vectors"]:
        Q = (H(v["q_x"]), H(v["q_y"]))
        r, s = H(v["r"]), H(v["s"])
        z = bytes.fromhex(v["m"])
        if c.low_s and s > c.n >> 1:
            keys.append(Q); hs.append(z); sigs.append((r, s)); exp.append(False)
            s = c.n - s
        keys.append(Q); hs.append(z); sigs.append((r, s)); exp.append(True)
        sb = bytearray(s.to_bytes(c.fb, "big"))
        sb[0] ^= 1
        keys.append(Q); hs.append(z); sigs.append((r, int.from_bytes(sb, "big"))); exp.append(False)
    assert eng.verify_prehash_batch(cname, keys, hs, sigs) == exp


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p224"])
def test_wycheproof_all_rows(eng, golden, cname):
    """All Wycheproof rows (k256/src/ecdsa.rs:345-424; new_wycheproof_test! for p256/p384).  DER parsing is
    host-side (strict, as the reference); every row that reaches arithmetic runs on the device."""
    c = o.curve(cname)
    blob = golden["wycheproof"][cname]
    hf = getattr(hashlib, blob["hash"])
    keys, hs, sigs, exp = [], [], [], []
    n_rows = 0
    for wx, wy, msg, sig, flag in blob["rows"]:
        n_rows += 1
        rs = o.der_parse_strict(bytes.fromhex(sig), c)
        if rs is None:
            assert not flag
            continue
        r, s = rs
        Q = (int.from_bytes(bytes.fromhex(wx)[-c.fb:], "big"), int.from_bytes(bytes.fromhex(wy)[-c.fb:], "big"))
        digest = hf(bytes.fromhex(msg)).digest()
        if c.low_s and 1 <= s < c.n and s > c.n >> 1:
            keys.append(Q); hs.append(digest); sigs.append((r, s)); exp.append(False)   # raw high-s: reject
            s = c.n - s                                                                 # runner normalises (ecdsa.rs:389)
        keys.append(Q); hs.append(digest); sigs.append((r, s)); exp.append(bool(flag))
    got = eng.verify_prehash_batch(cname, keys, hs, sigs)
    assert got == exp
    assert got == [o.verify_prehash(c, Q, h, r, s) for Q, h, (r, s) in zip(keys, hs, sigs)]
    assert sum(exp) >= 100 and n_rows == len(blob["rows"])


def test_prehash_length_cases(eng, golden):
    m = golden["misc"]["p256_prehash_sha384_verify"]
    assert eng.verify_prehash_batch("p256", [(H(m["qx"]), H(m["qy"]))], [bytes.fromhex(m["prehash"])], [(H(m["r"]), H(m["s"]))]) == [True]
    m = golden["misc"]["p384_prehash_sha256_verify"]
    assert eng.verify_prehash_batch("p384", [(H(m["qx"]), H(m["qy"]))], [bytes.fromhex(m["prehash"])], [(H(m["r"]), H(m["s"]))]) == [True]
    assert eng.verify_prehash_batch("p256", [o.P256.G], [b"\x01" * 15], [(1, 1)]) == [False]
    m = golden["misc"]["p256_rfc6979"]
    Q = o.mul_gen(o.P256, H(m["d"]))
    res = eng.verify_prehash_batch("p256", [Q, Q], [hashlib.sha256(t.encode()).digest() for t, _ in m["sigs"]],
                                   [(H(s[:64]), H(s[64:])) for _, s in m["sigs"]])
    assert res == [True, True]


def make_sigs(c, rng, n, nkeys=16):
    """Valid (Q, z, r, s) without private-key signing: R = aG + bQ, r = x(R) mod n, s = r/b, z = a*s."""
    keys = [o.mul_gen(c, rng.randrange(1, c.n)) for _ in range(nkeys)]
    rows = []
    for i in range(n):
        Q = keys[i % nkeys]
        a, b = rng.randrange(1, c.n), rng.randrange(1, c.n)
        R = o.pt_lincomb(c, [(c.G, a), (Q, b)])
        r = R[0] % c.n
        s = r * pow(b, -1, c.n) % c.n
        z = a * s % c.n
        if c.low_s and s > c.n >> 1:
            s = c.n - s          # R -> -R: same x, still valid
        rows.append([Q, z.to_bytes(c.fb, "big"), r, s])
    return rows


@pytest.mark.parametrize("cname", CUR)
def test_verify_synthetic_mask(eng, cname):
    """Synthetic batch with a deterministic 1/4 corrupted (bit flips in r, s, z, high-s twin, r >= n,
    off-curve key): mask must equal the oracle's."""
    c = o.curve(cname)
    rng = random.Random(71 + c.cid)
    n = 256 if c.fb == 32 else 96
    rows = make_sigs(c, rng, n)
    for i, row in enumerate(rows):
        if i % 4 != 1:
            continue
        kind = (i // 4) % 6
        if kind == 0:
            row[2] ^= 1 << rng.randrange(8 * c.fb - 2)
        elif kind == 1:
            row[3] ^= 1 << rng.randrange(8 * c.fb - 2)
        elif kind == 2:
            zb = bytearray(row[1]); zb[rng.randrange(c.fb)] ^= 0x10; row[1] = bytes(zb)
        elif kind == 3:
            row[3] = c.n - row[3]
        elif kind == 4:
            row[2] = c.n + 3 if c.n + 3 < 1 << 8 * c.fb else c.n
        else:
            row[0] = (row[0][0], row[0][1] ^ 1)
    keys = [r[0] for r in rows]; hs = [r[1] for r in rows]; sigs = [(r[2], r[3]) for r in rows]
    got = eng.verify_prehash_batch(cname, keys, hs, sigs)
    exp = [o.verify_prehash(c, Q, h, r, s) for Q, h, (r, s) in zip(keys, hs, sigs)]
    assert got == exp
    assert sum(got) >= n // 2 and sum(got) < n


@pytest.mark.parametrize("cname", CUR)
def test_verify_exceptional_cases(eng, cname):
    """Rows crafted so that the Jacobian fast path meets P + P, P + (-P) and identity accumulators (fixed-base
    part and window loop), plus GLV corner scalars; both verify kernels must give the oracle's answer."""
    from tests import crafted
    c = o.curve(cname)
    rows = crafted.exceptional_rows(c) + crafted.reduced_x_rows(c)
    keys = [r[0] for r in rows]; hs = [r[1] for r in rows]; sigs = [(r[2], r[3]) for r in rows]
    got = eng.verify_prehash_batch(cname, keys, hs, sigs)
    exp = [o.verify_prehash(c, Q, h, r, s) for Q, h, (r, s) in zip(keys, hs, sigs)]
    assert got == exp
    assert sum(got) > 5 and sum(got) < len(got)


def test_verify_kernels_agree(golden):
    """The complete-formula verify kernel (ECB200_VERIFY_V1=1) and the Jacobian fast path agree on Wycheproof."""
    import os
    import ecb200
    os.environ["ECB200_VERIFY_V1"] = "1"
    try:
        e1 = ecb200.Engine(0)
    finally:
        del os.environ["ECB200_VERIFY_V1"]
    e2 = ecb200.Engine(0)
    for cname in ("k256", "p256"):
        c = o.curve(cname)
        blob = golden["wycheproof"][cname]
        hf = getattr(hashlib, blob["hash"])
        keys, hs, sigs = [], [], []
        for wx, wy, msg, sig, flag in blob["rows"]:
            rs = o.der_parse_strict(bytes.fromhex(sig), c)
            if rs is None:
                continue
            keys.append((int.from_bytes(bytes.fromhex(wx)[-c.fb:], "big"), int.from_bytes(bytes.fromhex(wy)[-c.fb:], "big")))
            hs.append(hf(bytes.fromhex(msg)).digest())
            sigs.append(rs)
        assert e1.verify_prehash_batch(cname, keys, hs, sigs) == e2.verify_prehash_batch(cname, keys, hs, sigs)
    e1.close()
    e2.close()


@pytest.mark.parametrize("cname", ["p256", "p384", "sm2", "p192", "p224"])
def test_window_table_paths_agree(eng, cname):
    """Primeorder public-input path: batch-affine window tables (k_wintab, the default) against per-thread Jacobian tables
    (ECB200_WINTAB=0) on the same rows — signatures incl. corrupted ones and off-curve keys, P*k with identity / invalid /
    repeated points and projective inputs — at a size that gives every k_wintab thread several rows and a ragged tail."""
    import os
    import ecb200
    c = o.curve(cname)
    fb = c.fb
    os.environ["ECB200_WINTAB"] = "0"
    try:
        e0 = ecb200.Engine(0)
    finally:
        del os.environ["ECB200_WINTAB"]
    rng = random.Random(5 + c.cid)
    base = make_sigs(c, rng, 24)
    n = 5003                                            # > 3 rows per thread for part of the grid, not a multiple of anything
    keys, hs, sigs = [], [], []
    for i in range(n):
        Q, h, r, s = base[i % len(base)]
        if i % 7 == 3:
            s ^= 1 << (i % 200)
        if i % 11 == 5:
            Q = (Q[0], Q[1] ^ 1)                        # off-curve key: the table kernel substitutes G, the row is rejected
        if i % 13 == 6:
            Q = (0, 0)
        keys.append(Q); hs.append(h); sigs.append((r, s))
    a = eng.verify_prehash_batch(cname, keys, hs, sigs)
    b = e0.verify_prehash_batch(cname, keys, hs, sigs)
    assert a == b and 0 < sum(a) < n
    sample = list(range(0, n, 97))
    assert [a[i] for i in sample] == [o.verify_prehash(c, keys[i], hs[i], *sigs[i]) for i in sample]
    # P*k, public scalars
    m = 1200
    pts = [o.mul_gen(c, rng.randrange(1, c.n)) for _ in range(16)]
    pb = bytearray(b"".join(be(pts[i % 16], fb) for i in range(m)))
    kb = be([rng.randrange(c.n) if i % 17 else (0, 1, c.n - 1)[i % 3] for i in range(m)], fb)
    inf = bytearray(m)
    for i in range(0, m, 19):
        inf[i] = 1
    pb[2 * fb * 7: 2 * fb * 8] = be((1, 1), fb)         # off-curve
    out1, inv1 = eng.mul_batch(cname, bytes(pb), kb, bytes(inf), 0)
    out0, inv0 = e0.mul_batch(cname, bytes(pb), kb, bytes(inf), 0)
    assert out1 == out0 and inv1 == inv0 and inv1[7] == 1
    good = bytearray(pb)
    good[2 * fb * 7: 2 * fb * 8] = be(c.G, fb)          # oracle on the batch with the invalid row replaced; that row itself = identity slot
    exp = o.batch_mul_var_affine(c, bytes(good), bytes(inf), kb)
    st = len(out1) // m
    assert out1[:7 * st] == exp[:7 * st] and out1[8 * st:] == exp[8 * st:] and out1[7 * st:8 * st] == bytes(st)
    xyz = bytearray()
    for i in range(m):
        P = pts[i % 16]
        lam = rng.randrange(1, c.p)
        xyz += be((0, 5, 0), fb) if i % 23 == 4 else be((P[0] * lam % c.p, P[1] * lam % c.p, lam), fb)
    out1, _ = eng.mul_batch(cname, bytes(xyz), kb, None, 8)
    out0, _ = e0.mul_batch(cname, bytes(xyz), kb, None, 8)
    assert out1 == out0 == o.batch_mul_var_proj(c, bytes(xyz), kb)
    e0.close()


def test_large_batch_properties(eng):
    """Size-independent properties at a large size (2^18 here keeps the GPU test tier short; bench.py runs
    the full BASELINE sizes): (a) k*G via fixed-base == via variable-base with P = G; (b) (k1+k2)G ==
    lincomb; (c) verify accepts constructed signatures and rejects the tampered residue class; (d) results
    independent of chunking / position."""
    import ecb200
    c = o.K256
    n = 1 << 18
    rng = np.random.default_rng(0xB2000001)
    ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ks[:, 0] &= 0x7F
    kb = ks.tobytes()
    a = eng.mul_by_generator_batch("k256", kb, ecb200.FLAG_CT)
    gb = be(c.G, 32) * n
    b, inv = eng.mul_batch("k256", gb, kb, None, 0)
    assert a == b and not any(inv)
    # spot check against the oracle
    for i in list(range(0, n, n // 64))[:64]:
        assert a[33 * i:33 * i + 33] == o.slot_encode(c, o.mul_gen(c, int.from_bytes(kb[32 * i:32 * i + 32], "big")))
    # position independence: reversed batch gives reversed output
    rev = ks[::-1].copy().tobytes()
    ar = eng.mul_by_generator_batch("k256", rev, 0)
    assert np.array_equal(np.frombuffer(ar, np.uint8).reshape(n, 33)[::-1], np.frombuffer(a, np.uint8).reshape(n, 33))


def test_pinned_host_buffers(eng):
    """Host entry points DMA directly from / to page-locked caller buffers and stage pageable ones: same bytes."""
    import torch
    c = o.K256
    rng = random.Random(5)
    rows = make_sigs(c, rng, 300)
    for i in range(0, 300, 7):
        rows[i][3] ^= 2
    q = np.frombuffer(b"".join(be(r[0], 32) for r in rows), np.uint8).copy()
    z = np.frombuffer(b"".join(r[1] for r in rows), np.uint8).copy()
    rs = np.frombuffer(b"".join(be((r[2], r[3]), 32) for r in rows), np.uint8).copy()
    ref = eng.ecdsa_verify("k256", q.tobytes(), z.tobytes(), rs.tobytes())
    pq, pz, prs = (torch.from_numpy(a).pin_memory() for a in (q, z, rs))
    out = torch.zeros(300, dtype=torch.uint8).pin_memory()
    got = eng.ecdsa_verify("k256", pq.numpy(), pz.numpy(), prs.numpy(), out=out.numpy())
    assert bytes(got) == ref and 0 < sum(ref) < 300
    # mixed: pinned inputs, pageable output
    assert eng.ecdsa_verify("k256", pq.numpy(), z, prs.numpy()) == ref


FIRST_PIECE, PIECE = 227328, 5 * 227328     # abi.cu ECB200_FIRST_CHUNK / ECB200_CHUNK (wave-aligned pieces of the host pipeline)


@pytest.mark.parametrize("n", [FIRST_PIECE - 1, FIRST_PIECE + 1, FIRST_PIECE + PIECE + 5, FIRST_PIECE + 2 * PIECE, (1 << 18) + (1 << 20) + 5])
def test_pipeline_chunk_boundaries(eng, n):
    """Host pipeline pieces (227 328 then 1 136 640 rows): a periodic input must give the same periodic output at every
    position, whatever piece a row falls into."""
    c = o.K256
    per = 1009
    rng = random.Random(n)
    A = [rng.randrange(c.p) for _ in range(per)]
    B = [rng.randrange(c.p) for _ in range(per)]
    a = np.frombuffer(be(A, 32), np.uint8).reshape(per, 32)
    b = np.frombuffer(be(B, 32), np.uint8).reshape(per, 32)
    reps = (n + per - 1) // per
    aa = np.tile(a, (reps, 1))[:n].copy()
    bb = np.tile(b, (reps, 1))[:n].copy()
    out, ok = eng.field_op("k256", 0, 2, aa, bb)
    got = np.frombuffer(out, np.uint8).reshape(n, 32)
    exp = np.frombuffer(be([x * y % c.p for x, y in zip(A, B)], 32), np.uint8).reshape(per, 32)
    assert np.array_equal(got, np.tile(exp, (reps, 1))[:n])
    assert ok == b"\x01" * n


def test_abi_error_behaviour_and_reuse(eng):
    """Status codes are for misuse only (bad curve id, null buffers); empty batches succeed; scratch buffers regrow across
    calls of different sizes; two contexts on one device do not interfere."""
    import ctypes
    import ecb200
    lib = eng.lib
    buf = (ctypes.c_uint8 * 64)()
    assert lib.ecb200_mul_gen(eng.h, 7, 1, buf, buf, 0) == -1                       # unknown curve
    assert lib.ecb200_mul_gen(eng.h, 0, 1, None, buf, 0) == -1                      # null input
    assert lib.ecb200_ecdsa_verify(eng.h, 0, 1, buf, None, buf, buf) == -1
    assert lib.ecb200_decode_points(eng.h, 0, 1, buf, 8, 0, buf, buf) == -1         # stride too small
    assert lib.ecb200_decode_points(eng.h, 0, 1, buf, 33, 2, buf, buf) == -1        # unknown mode
    assert b"bad argument" in lib.ecb200_last_error(eng.h)
    # device-pointer calls index rows with int: more than 2^31 - 1 rows per call is refused, not truncated
    assert lib.ecb200_mul_gen_dev(eng.h, 0, 1 << 31, buf, buf, 0, None) == -1
    assert lib.ecb200_ecdsa_verify_dev(eng.h, 1, (1 << 31) + 5, buf, buf, buf, buf, None) == -1
    assert lib.ecb200_mul_gen(eng.h, 5, 1, None, buf, 0) == -1 and lib.ecb200_mul_gen(eng.h, 6, 1, buf, buf, 0) == -1   # last curve id is 5
    # n = 0 is fine, even with null buffers
    assert lib.ecb200_mul_gen(eng.h, 0, 0, None, None, 0) == 0
    assert lib.ecb200_ecdsa_verify(eng.h, 1, 0, None, None, None, None) == 0
    assert lib.ecb200_schnorr_verify(eng.h, 0, None, None, None, None) == 0
    assert eng.mul_by_generator_batch("k256", b"") == b""
    assert eng.ecdsa_verify("p256", b"", b"", b"") == b""
    assert eng.ecdsa_recover("k256", b"", b"", b"") == (b"", b"")
    assert eng.ecdsa_sign("p384", b"", b"", b"") == (b"", b"", b"")
    # growing then shrinking batches reuse / regrow the context's scratch
    c = o.K256
    e2 = ecb200.Engine(0)
    for n in (1, 1000, 17, 70000, 3):
        ks = be([(i * 7919 + 1) % c.n for i in range(n)], 32)
        a = eng.mul_by_generator_batch("k256", ks, ecb200.FLAG_CT)
        b = e2.mul_by_generator_batch("k256", ks, 0)
        assert a == b
        assert a[:33] == o.slot_encode(c, c.G)
    e2.close()


@pytest.mark.parametrize("cname,lg", [("k256", 22), ("p256", 22)])
def test_verify_full_size_mask(eng, cname, lg):
    """BASELINE.json's full batch size (configs[2] and configs[3]: 2^22 secp256k1 rows, 2^22 P-256 rows): the accept
    mask must equal the one implied by construction (1/16 of the rows corrupted in five ways), a sampled subset must agree
    with the oracle, and verifying the reversed batch must give the reversed mask (position independence)."""
    import ecb200
    wl = __import__("importlib").import_module("rustcrypto-elliptic-curves_b200.workloads")
    c = o.curve(cname)
    n = 1 << lg
    q, z, rs, exp = wl.make_verify_batch(wl.EngineBackend(eng, cname), cname, n, 0xB2000003 + c.cid)
    ok = np.frombuffer(eng.ecdsa_verify(cname, q, z, rs), np.uint8)
    assert np.array_equal(ok, exp)
    assert n - n // 16 <= int(exp.sum()) < n          # 1/16 corrupted; the high-s twins still verify on P-256
    for i in list(range(0, n, n // 97)) + [5, 21, 37, 53, 69]:
        Q = (int.from_bytes(q[i, :c.fb].tobytes(), "big"), int.from_bytes(q[i, c.fb:].tobytes(), "big"))
        r, s = int.from_bytes(rs[i, :c.fb].tobytes(), "big"), int.from_bytes(rs[i, c.fb:].tobytes(), "big")
        assert bool(ok[i]) == o.verify_prehashed(c, Q, z[i].tobytes(), r, s), i
    okr = np.frombuffer(eng.ecdsa_verify(cname, q[::-1].copy(), z[::-1].copy(), rs[::-1].copy()), np.uint8)
    assert np.array_equal(okr[::-1], ok)
