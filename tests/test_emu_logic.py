"""CPU-only logic check of the kernel bodies (host emulation of the PTX carry chains) against the
oracle.  This is NOT the parity gate (that is tests/test_gpu_parity.py on the B200); it keeps the
limb-level algorithms honest in the GPU-less container."""
import hashlib
import random

import pytest

from oracle import ecoracle as o
from tests import emu_lib

lib = emu_lib.load()
CUR = ["k256", "p256", "p384", "sm2", "p192", "p224"]


def field_op(c, which, op, a_list, b_list=None):
    fb = c.fb
    n = len(a_list)
    a = b"".join(v.to_bytes(fb, "big") for v in a_list)
    b = b"".join(v.to_bytes(fb, "big") for v in b_list) if b_list is not None else None
    out = emu_lib.buf(n * fb)
    ok = emu_lib.buf(n)
    assert lib.emu_field_op(c.cid, which, op, n, a, b, out, ok) == 0
    raw = bytes(out)
    return [int.from_bytes(raw[i * fb:(i + 1) * fb], "big") for i in range(n)], list(ok)


def edge_values(m):
    vals = [0, 1, 2, 3, m - 1, m - 2, m - 3, (m - 1) // 2, (m + 1) // 2, 2**32, 2**32 - 1, 2**32 + 977, 2**64 - 1,
            2**128, 2**128 - 1, 2**255 % m, (2**256 - 1) % m, 2**224 % m, (2**192 - 1) % m, 0xFFFFFFFF, m >> 32, m - 2**32]
    return sorted(set(v % m for v in vals))


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("which", [0, 1])
def test_field_ops(cname, which):
    c = o.curve(cname)
    m = c.p if which == 0 else c.n
    rng = random.Random(which * 10 + c.cid)
    ev = edge_values(m)
    A = [x for x in ev for _ in ev] + [rng.randrange(m) for _ in range(400)]
    B = [y for _ in ev for y in ev] + [rng.randrange(m) for _ in range(400)]
    for op, fn in ((0, lambda a, b: (a + b) % m), (1, lambda a, b: (a - b) % m), (2, lambda a, b: a * b % m)):
        got, ok = field_op(c, which, op, A, B)
        assert all(ok)
        exp = [fn(a, b) for a, b in zip(A, B)]
        assert got == exp, (cname, which, op, [i for i in range(len(A)) if got[i] != exp[i]][:3])
    got, _ = field_op(c, which, 3, A)
    assert got == [a * a % m for a in A]
    got, _ = field_op(c, which, 4, A)
    assert got == [(-a) % m for a in A]
    A2 = ev + [rng.randrange(m) for _ in range(20)]
    got, _ = field_op(c, which, 5, A2)
    assert got == [pow(a, m - 2, m) for a in A2]
    # the same inverse by divsteps (safegcd.cuh): edge values, powers of two, values around 2^30 limb boundaries
    A3 = A2 + [1 << k for k in range(0, 8 * c.fb - 1, 7)] + [(1 << k) - 1 for k in range(29, 8 * c.fb, 30)] + [m - (1 << k) for k in range(1, 8 * c.fb - 2, 29)]
    A3 = [a % m for a in A3] + [rng.randrange(m) for _ in range(300)]
    got, _ = field_op(c, which, 7, A3)
    assert got == [pow(a, m - 2, m) for a in A3]
    # non-canonical input is flagged
    _, ok = field_op(c, which, 0, [m, 2**(8 * c.fb) - 1], [0, 0])
    assert ok == [0, 0]


@pytest.mark.parametrize("cname", CUR)
def test_sqrt(cname):
    c = o.curve(cname)
    rng = random.Random(7)
    xs = [rng.randrange(c.p) for _ in range(12)]
    sq = [x * x % c.p for x in xs]
    got, ok = field_op(c, 0, 6, sq)
    assert all(ok)
    assert all(g * g % c.p == s for g, s in zip(got, sq))


def test_k256_golden_field(golden):
    c = o.K256
    k = golden["field"]["k256_risc0_8x32"]
    a, b = int(k["a"], 16), int(k["b"], 16)
    assert field_op(c, 0, 0, [a], [b])[0] == [int(k["add"], 16)]
    assert field_op(c, 0, 2, [a], [b])[0] == [int(k["mul"], 16)]
    assert field_op(c, 0, 3, [a])[0] == [int(k["square"], 16)]
    assert field_op(c, 0, 4, [a])[0] == [int(k["negate"], 16)]
    dbl = [int(h, 16) for h in golden["field"]["k256"]["dbl"]]
    assert field_op(c, 0, 0, dbl[:-1], dbl[:-1])[0] == dbl[1:]


def scalars_edge(c, rng, n):
    base = [0, 1, 2, c.n - 1, c.n - 2, c.n >> 1, (c.n >> 1) + 1, 1 << 128, (1 << 128) - 1, c.n, c.n + 1, 15, 16, 17]
    return base + [rng.randrange(c.n) for _ in range(n)]


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("ct", [0, 1])
def test_mul_var(cname, ct):
    c = o.curve(cname)
    rng = random.Random(11 + c.cid)
    ks = scalars_edge(c, rng, 6 if c.fb == 48 else 10)
    pts = [o.mul_gen(c, rng.randrange(1, c.n)) for _ in ks]
    pts[3] = c.G
    pts[4] = o.pt_neg(c, c.G)
    inf = [0] * len(ks)
    inf[5] = 1
    fb = c.fb
    pb = b"".join(P[0].to_bytes(fb, "big") + P[1].to_bytes(fb, "big") for P in pts)
    kb = b"".join((k % (1 << (8 * fb))).to_bytes(fb, "big") for k in ks)
    for compress in (0, 1):
        stride = 1 + (fb if compress else 2 * fb)
        out = emu_lib.buf(len(ks) * stride)
        invalid = emu_lib.buf(len(ks))
        assert lib.emu_mul_var(c.cid, ct, len(ks), 0, pb, bytes(inf), kb, out, compress, invalid, 3) == 0
        exp = o.batch_mul_var_affine(c, pb, bytes(inf), kb, bool(compress))
        assert bytes(out) == exp
        assert not any(invalid)
    if not ct:   # Jacobian fast path (affine inputs)
        stride = 1 + (fb if c.compress else 2 * fb)
        out = emu_lib.buf(len(ks) * stride)
        invalid = emu_lib.buf(len(ks))
        assert lib.emu_mul_var_fast(c.cid, len(ks), pb, bytes(inf), kb, out, int(c.compress), invalid) == 0
        assert bytes(out) == o.batch_mul_var_affine(c, pb, bytes(inf), kb)
        assert not any(invalid)
    # projective inputs with random Z, incl. Z = 0
    xyz = bytearray()
    for i, P in enumerate(pts):
        lam = rng.randrange(1, c.p)
        if i == 2:
            xyz += (0).to_bytes(fb, "big") + (7).to_bytes(fb, "big") + (0).to_bytes(fb, "big")
        else:
            xyz += (P[0] * lam % c.p).to_bytes(fb, "big") + (P[1] * lam % c.p).to_bytes(fb, "big") + lam.to_bytes(fb, "big")
    stride = 1 + (fb if c.compress else 2 * fb)
    out = emu_lib.buf(len(ks) * stride)
    assert lib.emu_mul_var(c.cid, ct, len(ks), 8, bytes(xyz), None, kb, out, int(c.compress), None, 0) == 0
    assert bytes(out) == o.batch_mul_var_proj(c, bytes(xyz), kb)
    # off-curve point is flagged invalid and yields the identity slot
    bad = (1).to_bytes(fb, "big") + (1).to_bytes(fb, "big")
    out = emu_lib.buf(stride)
    invalid = emu_lib.buf(1)
    lib.emu_mul_var(c.cid, ct, 1, 0, bad, None, (5).to_bytes(fb, "big"), out, int(c.compress), invalid, 0)
    assert list(invalid) == [1] and bytes(out) == bytes(stride)


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("ct", [0, 1])
def test_mul_gen(cname, ct, golden):
    c = o.curve(cname)
    rng = random.Random(3)
    ks = scalars_edge(c, rng, 12) + [int(k, 16) for k, _, _ in golden["group"].get(cname, {"mul": []})["mul"]]
    ks += [(1 << (8 * c.fb)) - 1, (1 << (8 * c.fb)) - 2, c.n - 8, int("8" * (2 * c.fb), 16), int("7" * (2 * c.fb), 16) % c.n]   # top-window carry
    fb = c.fb
    kb = b"".join((k % (1 << (8 * fb))).to_bytes(fb, "big") for k in ks)
    stride = 1 + (fb if c.compress else 2 * fb)
    out = emu_lib.buf(len(ks) * stride)
    assert lib.emu_mul_gen(c.cid, ct, len(ks), kb, out, int(c.compress)) == 0
    assert bytes(out) == o.batch_mul_gen(c, kb)


def split_gen_scalars(c, rng, W=5):
    """Scalars that stress the split fixed-base schedule: zero halves, one non-zero digit per half, runs of maximal digits
    (carries of the signed recoding through every window), the half boundary, the top window and its carry."""
    fb, n = c.fb, c.n
    bits = 8 * fb
    nwin = bits // W + 1
    nh = (nwin + 1) // 2
    cut = W * nh                                   # first bit of the upper half
    ks = scalars_edge(c, rng, 8)
    ks += [0, 1, 2, n - 1, n - 2, n, n + 1, (1 << bits) - 1, (1 << bits) - 2]
    ks += [1 << cut, (1 << cut) - 1, (1 << cut) + 1, (1 << (cut - 1)), (1 << (cut - 1)) - 1, ((1 << cut) - 1) ^ ((1 << (cut - W)) - 1)]
    ks += [rng.randrange(1 << cut), rng.randrange(1 << (bits - cut)) << cut]            # one half empty
    ks += [(1 << (W * w)) * v % n for w in (0, 1, nh - 1, nh, nh + 1, nwin - 2, nwin - 1) for v in (1, (1 << (W - 1)) - 1, 1 << (W - 1), (1 << (W - 1)) + 1, (1 << W) - 1)]
    ks += [int("8" * (2 * fb), 16) % n, int("7" * (2 * fb), 16) % n, int("F" * (2 * fb), 16) % n, n - 8, n >> 1, (n >> 1) + 1]
    half_pattern = sum(1 << (W * w + W - 1) for w in range(nwin - 1))                   # every signed window at -2^(W-1) + carry chain
    ks += [half_pattern % n, (half_pattern - 1) % n, (half_pattern + 1) % n, (n - half_pattern) % n]
    return ks


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("ct", [0, 1])
def test_mul_gen_split(cname, ct, golden):
    """split fixed-base path (two half sums with Jacobian mixed additions, complete addition in the normalisation)"""
    c = o.curve(cname)
    rng = random.Random(31 + ct)
    ks = split_gen_scalars(c, rng) + [rng.randrange(c.n) for _ in range(40)]
    ks += [int(k, 16) for k, _, _ in golden["group"].get(cname, {"mul": []})["mul"]]
    fb = c.fb
    kb = b"".join((k % (1 << (8 * fb))).to_bytes(fb, "big") for k in ks)
    stride = 1 + (fb if c.compress else 2 * fb)
    for nthreads in (0, 7):
        out = emu_lib.buf(len(ks) * stride)
        assert lib.emu_mul_gen2(c.cid, ct, len(ks), kb, out, int(c.compress), nthreads) == 0
        exp = o.batch_mul_gen(c, kb)
        got = bytes(out)
        bad = [i for i in range(len(ks)) if got[i * stride:(i + 1) * stride] != exp[i * stride:(i + 1) * stride]]
        assert not bad, (cname, ct, [hex(ks[i]) for i in bad[:4]])


@pytest.mark.parametrize("cname", CUR)
def test_batch_normalize(cname):
    c = o.curve(cname)
    rng = random.Random(21)
    fb = c.fb
    pts = []
    for i in range(37):
        if i in (0, 17, 36):
            pts.append((rng.randrange(c.p), rng.randrange(1, c.p), 0))
        else:
            P = o.mul_gen(c, rng.randrange(1, c.n))
            lam = rng.randrange(1, c.p)
            pts.append((P[0] * lam % c.p, P[1] * lam % c.p, lam))
    xyz = b"".join(v.to_bytes(fb, "big") for P in pts for v in P)
    exp = o.batch_normalize(c, pts)
    for nthreads in (1, 4, 0):
        xy = emu_lib.buf(len(pts) * 2 * fb)
        inf = emu_lib.buf(len(pts))
        assert lib.emu_batch_normalize(c.cid, len(pts), xyz, xy, inf, nthreads) == 0
        raw = bytes(xy)
        for i, P in enumerate(exp):
            if P is None:
                assert inf[i] == 1 and raw[i * 2 * fb:(i + 1) * 2 * fb] == bytes(2 * fb)
            else:
                assert inf[i] == 0
                assert raw[i * 2 * fb:(i + 1) * 2 * fb] == P[0].to_bytes(fb, "big") + P[1].to_bytes(fb, "big")


def run_verify(c, rows):
    """runs BOTH verify kernels (complete-formula v1 and the Jacobian fast path) and checks they agree"""
    fb = c.fb
    q = b"".join(r[0][0].to_bytes(fb, "big") + r[0][1].to_bytes(fb, "big") for r in rows)
    z = b"".join(r[1] for r in rows)
    rs = b"".join((r[2] % (1 << 8 * fb)).to_bytes(fb, "big") + (r[3] % (1 << 8 * fb)).to_bytes(fb, "big") for r in rows)
    ok = emu_lib.buf(len(rows))
    assert lib.emu_verify(c.cid, len(rows), q, z, rs, ok) == 0
    ok2 = emu_lib.buf(len(rows))
    assert lib.emu_verify2(c.cid, len(rows), q, z, rs, ok2, 3) == 0
    assert list(ok) == list(ok2)
    # third path: rows grouped by public key, per-key window tables, no doublings - fed in one chunk, in three chunks (groups
    # carried over) and one row at a time; the distinct-key count must match
    n = len(rows)
    distinct = len({r[0] for r in rows})
    for chunk in sorted({n, max(1, (n + 2) // 3), 1 if n <= 40 else 7}):
        ok3 = emu_lib.buf(n)
        assert lib.emu_verify_keytab(c.cid, 0, n, q, z, rs, ok3, chunk, n, 2) == distinct
        assert list(ok3) == list(ok2), chunk
    return list(ok2), o.batch_verify(c, q, z, rs)


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p224"])
def test_verify_wycheproof_sample(golden, cname):
    c = o.curve(cname)
    blob = golden["wycheproof"][cname]
    hf = getattr(hashlib, blob["hash"])
    rows = []
    for wx, wy, msg, sig, flag in blob["rows"]:
        rsv = o.der_parse_strict(bytes.fromhex(sig), c)
        if rsv is None:
            continue
        r, s = rsv
        if r >> (8 * c.fb) or s >> (8 * c.fb):
            continue
        wxb, wyb = bytes.fromhex(wx)[-c.fb:], bytes.fromhex(wy)[-c.fb:]
        Q = (int.from_bytes(wxb, "big"), int.from_bytes(wyb, "big"))
        zb = o.bits2field(c, hf(bytes.fromhex(msg)).digest())
        rows.append((Q, zb, r, s))
        if c.low_s and s > c.n >> 1 and 1 <= s < c.n:
            rows.append((Q, zb, r, c.n - s))
    rows = rows[::3] if cname == "p384" else rows[::2]
    got, exp = run_verify(c, rows)
    assert got == list(exp)
    assert sum(got) > 20


@pytest.mark.parametrize("cname", CUR)
def test_verify_synthetic(cname):
    c = o.curve(cname)
    rng = random.Random(77)
    rows = []
    for i in range(10):
        d = rng.randrange(1, c.n)
        Q = o.mul_gen(c, d)
        z = rng.randrange(1 << (8 * c.fb)).to_bytes(c.fb, "big")
        k = rng.randrange(1, c.n)
        r = o.mul_gen(c, k)[0] % c.n
        s = pow(k, -1, c.n) * (o.reduce_once(c, int.from_bytes(z, "big")) + r * d) % c.n
        if c.low_s and s > c.n >> 1:
            s = c.n - s
        rows.append((Q, z, r, s))
        if i % 3 == 0:
            rows.append((Q, z, r, c.n - s))          # high-s twin: reject on k256, accept elsewhere
        if i % 4 == 0:
            rows.append((Q, z, r ^ 1, s))
            rows.append((Q, z, 0, s))
            rows.append((Q, z, r, c.n))
            rows.append(((Q[0], Q[1] ^ 1), z, r, s))  # off-curve key
    got, exp = run_verify(c, rows)
    assert got == list(exp)
    assert 0 < sum(got) < len(got)


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p224"])
@pytest.mark.parametrize("gw", [4, 7, 8])
def test_verify_fixed_base_window_widths(cname, gw, monkeypatch):
    """u1*G from the big fixed-base table with windows of any width (jac.cuh add_fixed_base): word-aligned (4, 8) and
    straddling (7; the other emulation tests run at 5) widths, top windows shorter than the rest, u1 with all-ones words
    (carries through every window of the recoding), u1 = 0 and 1."""
    monkeypatch.setenv("ECB_EMU_GW", str(gw))
    c = o.curve(cname)
    rng = random.Random(900 + gw)
    rows = []
    top = 1 << (8 * c.fb)
    for u1 in [0, 1, c.n - 1, (top >> 1) % c.n, (top - 1) % c.n, 0x88888888, rng.randrange(c.n), rng.randrange(c.n)]:
        # a signature whose u1 = z / s is the chosen value: k = u1 + u2 d with u2 = r / s  =>  s = r d / (k - u1), z = u1 s
        d = rng.randrange(1, c.n)
        Q = o.mul_gen(c, d)
        k = rng.randrange(1, c.n)
        if (k - u1) % c.n == 0:
            k = k % (c.n - 1) + 1
        r = o.mul_gen(c, k)[0] % c.n
        s = r * d * pow((k - u1) % c.n, -1, c.n) % c.n
        z = u1 * s % c.n
        if r == 0 or s == 0:
            continue
        if c.low_s and s > c.n >> 1:          # (r, n - s) verifies with -u1, -u2: same x(R)
            s = c.n - s
        rows.append((Q, z.to_bytes(c.fb, "big"), r, s))
        rows.append((Q, z.to_bytes(c.fb, "big"), r, s % (c.n - 1) + 1))
    got, exp = run_verify(c, rows)
    assert got == list(exp)
    assert sum(got) >= len(rows) // 2 - 1


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("gw", [5, 8])
def test_verify_fixed_base_window_collisions(cname, gw, monkeypatch):
    """Rows built so that the accumulator u2*Q + (windows of u1*G below j) equals, or is the negative of, the table entry of
    window j (tests/crafted.py fixed_base_collision_rows): the P + P and P + (-P) branches of the gathered mixed addition at
    EVERY window position of the recoding incl. the unsigned top window, on the per-row and the per-key-table kernels."""
    from tests import crafted
    monkeypatch.setenv("ECB_EMU_GW", str(gw))
    c = o.curve(cname)
    nwin = (8 * c.fb + gw - 1) // gw
    windows = list(range(nwin)) if c.fb <= 32 and gw == 8 else sorted(set(list(range(0, nwin, 5 if c.fb <= 32 else 9)) + [1, nwin - 2, nwin - 1]))
    rows = crafted.fixed_base_collision_rows(c, gw, windows)
    assert len(rows) >= 2 * len(windows) - 2
    # twins with a different r must be rejected on the same exceptional path
    rows += [(Q, z, r % (c.n - 1) + 1, s) for Q, z, r, s in rows[::3]]
    got, exp = run_verify(c, rows)
    assert got == list(exp)
    assert sum(got) >= 2 * len(windows) - 4 and sum(got) < len(got)


@pytest.mark.parametrize("cname", ["k256", "p256", "sm2", "p224"])
def test_verify_keytab_reused_keys_and_overflow(cname):
    """Few keys, many rows (the shape the per-key tables exist for), incl. an off-curve key shared by several rows, corrupted
    signatures, rows whose digits are all zero in one GLV half, and a table capacity smaller than the number of keys."""
    c = o.curve(cname)
    fb = c.fb
    rng = random.Random(5150 + c.cid)
    keys = [rng.randrange(1, c.n) for _ in range(5)]
    rows = []
    for i in range(60):
        d = keys[i % 5]
        Q = o.mul_gen(c, d)
        z = rng.randrange(1 << (8 * fb)).to_bytes(fb, "big")
        k = rng.randrange(1, c.n)
        r = o.mul_gen(c, k)[0] % c.n
        s = pow(k, -1, c.n) * (o.reduce_once(c, int.from_bytes(z, "big")) + r * d) % c.n
        if c.low_s and s > c.n >> 1:
            s = c.n - s
        if i % 7 == 3:
            s ^= 4
        if i % 11 == 5:
            Q = (Q[0], Q[1] ^ 1)                       # the same broken key bytes recur: one invalid group
        rows.append((Q, z, r, s))
    got, exp = run_verify(c, rows)
    assert got == list(exp) and 20 < sum(got) < 60
    q = b"".join(r[0][0].to_bytes(fb, "big") + r[0][1].to_bytes(fb, "big") for r in rows)
    z = b"".join(r[1] for r in rows)
    rs = b"".join(r[2].to_bytes(fb, "big") + r[3].to_bytes(fb, "big") for r in rows)
    ok = emu_lib.buf(60)
    assert lib.emu_verify_keytab(c.cid, 0, 60, q, z, rs, ok, 60, 3, 2) < -3      # more distinct keys than capacity: reported, not mis-verified


@pytest.mark.parametrize("kt_w", [4, 5, 6])
def test_verify_keytab_window_widths(kt_w):
    """The per-key tables with 4-, 5- and 6-bit windows (ECB_KT_W): recoding across word boundaries, the unsigned top window,
    the table-construction tree of depth W - 1 - on every curve family (GLV halves, 6-, 7-, 8- and 12-limb scalars), with crafted
    scalars whose windows are all-maximal / all-minimal, against the oracle."""
    from tests import crafted
    vlib = emu_lib.load_variant(kt_w)
    for cname in CUR:
        c = o.curve(cname)
        fb = c.fb
        rng = random.Random(900 + kt_w + c.cid)
        rows = []
        keys = [rng.randrange(1, c.n) for _ in range(3)]
        for i in range(14 if fb <= 32 else 8):
            d = keys[i % 3]
            Q = o.mul_gen(c, d)
            z = rng.randrange(1 << (8 * fb)).to_bytes(fb, "big")
            k = rng.randrange(1, c.n)
            r = o.mul_gen(c, k)[0] % c.n
            s = pow(k, -1, c.n) * (o.reduce_once(c, int.from_bytes(z, "big")) + r * d) % c.n
            if c.low_s and s > c.n >> 1:
                s = c.n - s
            if i % 5 == 2:
                r ^= 2
            rows.append((Q, z, r, s))
        rows += crafted.exceptional_rows(c)[::4] + crafted.reduced_x_rows(c)[:4]
        # u2 = r / s with every window at its extreme: choose s = r / u2 for patterned u2 (verdict comes from the oracle)
        for pat in ("f" * (2 * fb), "8" * (2 * fb), "7" * (2 * fb), "10" * fb, "0f" * fb):
            u2 = int(pat, 16) % c.n or 1
            r = rng.randrange(1, c.n)
            rows.append((o.mul_gen(c, keys[0]), rng.randrange(1 << (8 * fb)).to_bytes(fb, "big"), r, r * pow(u2, -1, c.n) % c.n or 1))
        q = b"".join(x[0][0].to_bytes(fb, "big") + x[0][1].to_bytes(fb, "big") for x in rows)
        z = b"".join(x[1] for x in rows)
        rs = b"".join((x[2] % (1 << 8 * fb)).to_bytes(fb, "big") + (x[3] % (1 << 8 * fb)).to_bytes(fb, "big") for x in rows)
        n = len(rows)
        ok = emu_lib.buf(n)
        assert vlib.emu_verify_keytab(c.cid, 0, n, q, z, rs, ok, max(1, n // 2), n, 2) == len({x[0] for x in rows})
        assert bytes(ok) == o.batch_verify(c, q, z, rs), (cname, kt_w)
        assert 0 < sum(ok) < n


@pytest.mark.parametrize("cname", CUR)
def test_verify_exceptional_cases(cname):
    """P + P, P + (-P) and identity accumulators inside the Jacobian fast path, GLV corner scalars."""
    from tests import crafted
    c = o.curve(cname)
    rows = crafted.exceptional_rows(c)
    if c.fb == 48:
        rows = rows[::3]
    rows += crafted.reduced_x_rows(c)          # x(R) >= n: the second candidate r + n of the inversion-free comparison
    got, exp = run_verify(c, rows)
    assert got == list(exp)
    assert sum(got) > 5 and sum(got) < len(got)


@pytest.mark.parametrize("cname", CUR)
@pytest.mark.parametrize("which", [0, 1])
def test_field_mul_structured_limbs(cname, which):
    """Operands built from extreme 32-bit limbs (0, 1, 2^32-1, 2^32-2, 2^31, random): exercises every carry /
    borrow path of the reductions (special-form fold for k256, shifted-addition Montgomery for the sparse primes)."""
    c = o.curve(cname)
    m = c.p if which == 0 else c.n
    L = c.fb // 4
    rng = random.Random(99 + which + 7 * c.cid)
    pool = [0, 1, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF, 2]

    def val():
        v = 0
        for i in range(L):
            limb = rng.choice(pool) if rng.random() < 0.8 else rng.getrandbits(32)
            v |= limb << (32 * i)
        return v % m

    A = [val() for _ in range(3000)]
    B = [val() for _ in range(3000)]
    got, ok = field_op(c, which, 2, A, B)
    assert all(ok)
    assert got == [a * b % m for a, b in zip(A, B)]
    got, _ = field_op(c, which, 3, A)
    assert got == [a * a % m for a in A]
    got, _ = field_op(c, which, 0, A, B)
    assert got == [(a + b) % m for a, b in zip(A, B)]
    got, _ = field_op(c, which, 1, A, B)
    assert got == [(a - b) % m for a, b in zip(A, B)]


# ---------------------------------------------------------------------------------------------
# SURVEY §8f rows on the host emulation (logic only; parity is asserted on the B200)
from tests import nextrows  # noqa: E402

VM_ECDSA, VM_SM2DSA, VM_SCHNORR, VM_RECOVER = 0, 1, 2, 3


@pytest.mark.parametrize("cname", CUR)
def test_decode_points(cname):
    c = o.curve(cname)
    slots, stride, st, xy = nextrows.decode_cases(c, n_random=4 if c.fb == 48 else 8)
    n = len(st)
    got_xy, got_st = emu_lib.buf(n * 2 * c.fb), emu_lib.buf(n)
    assert lib.emu_decode(c.cid, n, 0, slots, stride, got_xy, got_st) == 0
    assert bytes(got_st) == st and bytes(got_xy) == xy
    enc, st, xy = nextrows.compact_cases(c, n_random=3)
    n = len(st)
    got_xy, got_st = emu_lib.buf(n * 2 * c.fb), emu_lib.buf(n)
    assert lib.emu_decode(c.cid, n, 1, enc, c.fb, got_xy, got_st) == 0
    assert bytes(got_st) == st and bytes(got_xy) == xy


@pytest.mark.parametrize("cname", ["k256", "p256"])
def test_recover(cname, golden):
    c = o.curve(cname)
    zb, rsb, ids, exp_keys, exp_ok = nextrows.recover_cases(c, golden, n_random=5)
    n = len(ids)
    stride = len(exp_keys) // n
    ok, keys = emu_lib.buf(n), emu_lib.buf(n * stride)
    assert lib.emu_verify_mode(c.cid, VM_RECOVER, n, None, zb, rsb, ids, ok, keys, int(c.compress), 3) == 0
    assert bytes(ok) == exp_ok
    assert bytes(keys) == exp_keys
    assert 0 < sum(exp_ok) < n


def test_schnorr(golden):
    pkb, eb, sb, exp = nextrows.schnorr_cases(golden, n_random=4)
    n = len(exp)
    ok = emu_lib.buf(n)
    assert lib.emu_verify_mode(0, VM_SCHNORR, n, pkb, eb, sb, None, ok, None, 0, 2) == 0
    assert bytes(ok) == exp
    assert 0 < sum(exp) < n


def test_sm2dsa(golden):
    qb, eb, rsb, exp = nextrows.sm2dsa_cases(golden, n_random=5)
    n = len(exp)
    ok = emu_lib.buf(n)
    assert lib.emu_verify_mode(3, VM_SM2DSA, n, qb, eb, rsb, None, ok, None, 0, 2) == 0
    assert bytes(ok) == exp
    assert 0 < sum(exp) < n
    ok3 = emu_lib.buf(n)                               # the same rows through the per-key-table path
    assert lib.emu_verify_keytab(3, VM_SM2DSA, n, qb, eb, rsb, ok3, 4, n, 2) > 0
    assert bytes(ok3) == exp


@pytest.mark.parametrize("cname", ["k256", "p256"])
def test_sign(cname, golden):
    c = o.curve(cname)
    db, kb, zb, rs, rid, okx = nextrows.sign_cases(c, golden, n_random=4)
    n = len(okx)
    g_rs, g_rid, g_ok = emu_lib.buf(n * 2 * c.fb), emu_lib.buf(n), emu_lib.buf(n)
    assert lib.emu_sign(c.cid, n, db, kb, zb, g_rs, g_rid, g_ok, 2) == 0
    assert bytes(g_ok) == okx and bytes(g_rs) == rs and bytes(g_rid) == rid
