"""GPU parity tests for the SURVEY §8f rows — SEC1 decoding, verification with encoded keys, public-key recovery,
BIP340 Schnorr, SM2DSA, signing — through the C ABI against the oracle and the reference's own vectors."""
import random

import numpy as np
import pytest

from oracle import ecoracle as o
from tests import nextrows

pytestmark = pytest.mark.gpu
CUR = ["k256", "p256", "p384", "sm2", "p192", "p224"]


@pytest.fixture(scope="module")
def eng():
    import ecb200
    e = ecb200.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("cname", CUR)
def test_decode_points(eng, cname):
    c = o.curve(cname)
    slots, stride, st, xy = nextrows.decode_cases(c, n_random=40)
    got_xy, got_st = eng.decode_points(cname, slots, stride, 0)
    assert got_st == st and got_xy == xy
    enc, st, xy = nextrows.compact_cases(c, n_random=20)
    got_xy, got_st = eng.decode_points(cname, enc, c.fb, 1)
    assert got_st == st and got_xy == xy
    # compressed-only stride
    rng = random.Random(2)
    pts = [o.mul_gen(c, rng.randrange(1, c.n)) for _ in range(33)]
    enc = b"".join(o.sec1_encode(c, P, True) for P in pts)
    got_xy, got_st = eng.decode_points(cname, enc, 1 + c.fb, 0)
    assert got_st == b"\x01" * 33
    assert got_xy == b"".join(nextrows.be(P[0], c.fb) + nextrows.be(P[1], c.fb) for P in pts)
    assert eng.decode_points(cname, b"", 1 + c.fb, 0) == (b"", b"")


@pytest.mark.parametrize("cname", CUR)
def test_verify_with_sec1_keys(eng, cname):
    """from_sec1_bytes + verify_prehash: compressed and uncompressed keys in one batch, bad encodings rejected."""
    c = o.curve(cname)
    rng = random.Random(31 + c.cid)
    fb = c.fb
    stride = 1 + 2 * fb
    keys, zs, rss, exp = [], [], [], []
    for i in range(48):
        d, k, z, (r, s, _) = nextrows.make_sig(c, rng)
        Q = o.mul_gen(c, d)
        enc = o.sec1_encode(c, Q, i % 2 == 0)
        good = True
        if i % 6 == 3:
            enc = bytes([enc[0] ^ 1]) + enc[1:] if enc[0] in (2, 3) else b"\x04" + enc[1:-1] + bytes([enc[-1] ^ 1])
            good = False                                       # other root / off-curve point
        if i % 8 == 5:
            enc = bytes(1)                                      # identity key: VerifyingKey rejects it
            good = False
        if i % 12 == 7:
            s ^= 4
            good = False
        keys.append(enc + bytes(stride - len(enc))); zs.append(z); rss.append(nextrows.be(r, fb) + nextrows.be(s, fb))
        exp.append(1 if good else 0)
    ok = eng.ecdsa_verify_sec1(cname, b"".join(keys), stride, b"".join(zs), b"".join(rss))
    # oracle: decode, then verify
    ref = []
    for enc, z, rs in zip(keys, zs, rss):
        raw = enc.rstrip(b"\x00") if enc[0] == 0 else enc[:1 + (fb if enc[0] in (2, 3) else 2 * fb)]
        okd, Q = o.sec1_decode(c, raw if raw else b"\x00")
        ref.append(1 if okd and Q is not None and o.verify_prehashed(c, Q, z, int.from_bytes(rs[:fb], "big"), int.from_bytes(rs[fb:], "big")) else 0)
    assert list(ok) == ref
    assert list(ok) == exp
    assert 0 < sum(ok) < len(ok)


@pytest.mark.parametrize("cname", CUR)
def test_recover(eng, cname, golden):
    c = o.curve(cname)
    zb, rsb, ids, exp_keys, exp_ok = nextrows.recover_cases(c, golden, n_random=40 if c.fb == 32 else 16)
    keys, ok = eng.ecdsa_recover(cname, zb, rsb, ids)
    assert ok == exp_ok
    assert keys == exp_keys
    assert 0 < sum(ok) < len(ok)
    if cname == "k256":   # reference vectors sit at the tail of the batch (k256/src/ecdsa.rs:278-343)
        vs = golden["next"]["k256_recovery"]["vectors"]
        n = len(ids)
        for j, v in enumerate(vs):
            i = n - len(vs) - 1 + j
            assert keys[33 * i:33 * i + 33].hex() == v["pk"] and ok[i] == 1


def test_schnorr_bip340(eng, golden):
    pkb, eb, sb, exp = nextrows.schnorr_cases(golden, n_random=64)
    ok = eng.schnorr_verify(pkb, eb, sb)
    assert ok == exp
    # the 15 BIP340 vectors are rows 0..14 (k256/src/schnorr.rs:217-449)
    want = [True] * 4 + [v["valid"] for v in golden["next"]["bip340_verify"]["vectors"]]
    assert [bool(x) for x in ok[:15]] == want
    assert 0 < sum(ok) < len(ok)


def test_sm2dsa(eng, golden):
    qb, eb, rsb, exp = nextrows.sm2dsa_cases(golden, n_random=48)
    ok = eng.sm2dsa_verify(qb, eb, rsb)
    assert ok == exp
    assert ok[0] == 1          # sm2/tests/sm2dsa.rs:16-32
    assert 0 < sum(ok) < len(ok)


@pytest.mark.parametrize("cname", CUR)
def test_sign(eng, cname, golden):
    c = o.curve(cname)
    db, kb, zb, rs, rid, okx = nextrows.sign_cases(c, golden, n_random=40 if c.fb == 32 else 12)
    g_rs, g_rid, g_ok = eng.ecdsa_sign(cname, db, kb, zb)
    assert g_ok == okx and g_rs == rs and g_rid == rid
    # sign -> verify -> recover round trip on the device
    n = len(okx)
    fb = c.fb
    good = [i for i in range(n) if okx[i]]
    qs = b"".join(nextrows.be(v, fb) for i in good for v in o.mul_gen(c, int.from_bytes(db[fb * i:fb * i + fb], "big")))
    z2 = b"".join(zb[fb * i:fb * i + fb] for i in good)
    rs2 = b"".join(g_rs[2 * fb * i:2 * fb * i + 2 * fb] for i in good)
    assert eng.ecdsa_verify(cname, qs, z2, rs2) == b"\x01" * len(good)
    keys, ok = eng.ecdsa_recover(cname, z2, rs2, bytes(g_rid[i] for i in good), 4)   # uncompressed slots
    assert ok == b"\x01" * len(good)
    assert b"".join(keys[(1 + 2 * fb) * j + 1:(1 + 2 * fb) * (j + 1)] for j in range(len(good))) == qs


def test_next_rows_large_batch_properties(eng):
    """2^16 rows: sign -> verify accepts everything, recover returns the signing key, compressed keys verify too;
    Schnorr verifies a tiled batch identically at every position."""
    import ecb200
    c = o.K256
    n = 1 << 16
    rng = np.random.default_rng(0xB2000007)
    def scal():
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 0] &= 0x7F
        a[:, 31] |= 1
        return a
    d, k, z = scal(), scal(), rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    rs, rid, ok = eng.ecdsa_sign("k256", d.tobytes(), k.tobytes(), z.tobytes())
    assert ok == b"\x01" * n
    pub = eng.mul_by_generator_batch("k256", d.tobytes(), ecb200.FLAG_CT)           # compressed SEC1 keys
    assert eng.ecdsa_verify_sec1("k256", pub, 33, z.tobytes(), rs) == b"\x01" * n
    keys, ok = eng.ecdsa_recover("k256", z.tobytes(), rs, rid)
    assert ok == b"\x01" * n and keys == pub
    # tamper: every row's s low bit
    bad = np.frombuffer(rs, np.uint8).reshape(n, 64).copy()
    bad[:, 63] ^= 1
    assert sum(eng.ecdsa_verify_sec1("k256", pub, 33, z.tobytes(), bad.tobytes())) == 0
    for i in (0, 1, n // 2, n - 1):   # oracle spot checks
        di, ki = int.from_bytes(d[i].tobytes(), "big"), int.from_bytes(k[i].tobytes(), "big")
        r, s, recid = o.sign_prehashed(c, di, ki, z[i].tobytes())
        assert rs[64 * i:64 * i + 64] == nextrows.be(r, 32) + nextrows.be(s, 32) and rid[i] == recid


def test_dev_pointer_variants(eng, golden):
    """_dev entry points (device buffers + stream) of the rows not exercised by bench.py: decode and SM2DSA."""
    import torch
    dev = torch.device("cuda:0")
    st = torch.cuda.Stream(device=dev)
    c = o.SM2
    qb, eb, rsb, exp = nextrows.sm2dsa_cases(golden, n_random=16)
    n = len(exp)
    t = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
    d_ok = torch.zeros(n, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(st):
        eng.sm2dsa_verify_dev(n, t(qb), t(eb), t(rsb), d_ok, st.cuda_stream)
    st.synchronize()
    assert bytes(d_ok.cpu().numpy().tobytes()) == exp
    for cname in ("k256", "p384"):
        c = o.curve(cname)
        slots, stride, status, xy = nextrows.decode_cases(c, n_random=10)
        n = len(status)
        d_xy = torch.zeros(n * 2 * c.fb, dtype=torch.uint8, device=dev)
        d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
        with torch.cuda.stream(st):
            eng.decode_points_dev(cname, n, t(slots), stride, 0, d_xy, d_st, st.cuda_stream)
        st.synchronize()
        assert bytes(d_st.cpu().numpy().tobytes()) == status and bytes(d_xy.cpu().numpy().tobytes()) == xy


@pytest.mark.parametrize("gw", [4, 8])
def test_fixed_base_window_widths(gw):
    """ECB200_GW selects the window width of the big fixed-base table (signed windows, unsigned top window): every width
    gives the oracle's answers on the crafted rows, incl. top-window carries."""
    import os
    import ecb200
    from tests import crafted
    os.environ["ECB200_GW"] = str(gw)
    try:
        e = ecb200.Engine(0)
    finally:
        del os.environ["ECB200_GW"]
    for cname in ("k256", "p256"):
        c = o.curve(cname)
        rows = crafted.exceptional_rows(c) + crafted.reduced_x_rows(c)
        # u1 with all-ones / alternating top windows: z chosen so that u1 = z / s hits them (s = 1 => u1 = z)
        rng = random.Random(gw)
        for u1 in ((1 << (8 * c.fb)) - 1, c.n - 1, int("8" * (2 * c.fb), 16) % c.n, int("7F" * c.fb, 16), 1 << (8 * c.fb - 1)):
            row = crafted.craft(c, u1 % c.n, rng.randrange(1, c.n), rng.randrange(1, c.n))
            if row:
                rows.append(row)
        keys = [r[0] for r in rows]; hs = [r[1] for r in rows]; sigs = [(r[2], r[3]) for r in rows]
        got = e.verify_prehash_batch(cname, keys, hs, sigs)
        assert got == [o.verify_prehash(c, Q, h, r, s) for Q, h, (r, s) in zip(keys, hs, sigs)]
        assert sum(got) > 5
    e.close()


def test_plain_c_client(tmp_path):
    """A plain C program compiled against include/ecb200.h and linked with libecb200.so (no Python in the loop)."""
    import os
    import subprocess
    import ecb200
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cabi_client")
    libdir = os.path.dirname(ecb200.LIB_PATH)
    subprocess.check_call(["gcc", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cabi", "client.c"), "-o", exe,
                           "-L", libdir, "-lecb200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "cabi client ok" in out.stdout
