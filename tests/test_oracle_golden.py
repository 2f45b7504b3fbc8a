"""Pins the oracle (oracle/ecoracle.py) against every golden vector the reference holds for the
hot path.  CPU only."""
import hashlib

import pytest

from oracle import ecoracle as o

H = lambda s: int(s, 16)


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p192", "p224"])
def test_add_vectors(golden, cname):
    # ADD_TEST_VECTORS[i] == (i+1)*G  (k256 projective.rs:859-967, primeorder/src/dev.rs:7-157)
    c = o.curve(cname)
    acc = None
    for i, (x, y) in enumerate(golden["group"][cname]["add"]):
        acc = o.pt_add(c, acc, c.G)
        assert acc == (H(x), H(y)), i
        assert o.mul_gen(c, i + 1) == acc
    # add vs double (primeorder/src/dev.rs:107-111)
    assert o.pt_add(c, c.G, c.G) == o.pt_dbl(c, c.G)
    assert o.pt_add(c, c.G, o.pt_neg(c, c.G)) is None


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p192", "p224"])
def test_mul_vectors(golden, cname):
    c = o.curve(cname)
    for k, x, y in golden["group"][cname]["mul"]:
        assert o.mul_gen(c, H(k)) == (H(x), H(y))
    if cname == "k256":
        for k, x, y in golden["group"][cname]["mul"][:8]:
            assert o.k256_lincomb_reference_algorithm([(c.G, H(k))]) == (H(x), H(y))


@pytest.mark.parametrize("cname", ["k256", "p256"])
def test_field_dbl_vectors(golden, cname):
    c = o.curve(cname)
    v = 1
    for hx in golden["field"][cname]["dbl"]:
        assert v == H(hx)
        v = v * 2 % c.p


def test_risc0_8x32_kats(golden):
    k = golden["field"]["k256_risc0_8x32"]
    p = o.K256.p
    a, b = H(k["a"]), H(k["b"])
    assert (a + b) % p == H(k["add"])
    assert (-a - b) % p == H(k["add_negated"])
    assert (-a) % p == H(k["negate"])
    assert a * b % p == H(k["mul"])
    assert a * a % p == H(k["square"])


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p192", "p224"])
def test_ecdsa_kats(golden, cname):
    # ecdsa_core::new_verification_test!: verify OK; flip bit 0 of s[0] => Err (p256/src/ecdsa.rs:184-192)
    c = o.curve(cname)
    vs = golden["ecdsa"][cname]["vectors"]
    assert vs
    for v in vs:
        Q = (H(v["q_x"]), H(v["q_y"]))
        assert o.mul_gen(c, H(v["d"])) == Q
        z = bytes.fromhex(v["m"])
        r, s = H(v["r"]), H(v["s"])
        if c.low_s and s > c.n >> 1:
            assert not o.verify_prehash(c, Q, z, r, s)
            s = c.n - s
        assert o.verify_prehash(c, Q, z, r, s)
        sb = bytearray(bytes.fromhex(v["s"]))
        sb[0] ^= 1
        assert not o.verify_prehash(c, Q, z, r, int.from_bytes(sb, "big"))


def _wx(c, hx):
    b = bytes.fromhex(hx)
    if len(b) >= c.fb:
        assert not any(b[:len(b) - c.fb])
        b = b[len(b) - c.fb:]
    return int.from_bytes(b, "big")


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p224"])
def test_wycheproof(golden, cname):
    # runners: k256/src/ecdsa.rs:345-424 (normalises s), ecdsa_core::new_wycheproof_test! (p256/p384)
    c = o.curve(cname)
    blob = golden["wycheproof"][cname]
    hf = getattr(hashlib, blob["hash"])
    n_der = n_range = n_math = 0
    for i, (wx, wy, msg, sig, flag) in enumerate(blob["rows"]):
        Q = (_wx(c, wx), _wx(c, wy))
        assert o.on_curve(c, *Q)
        rs = o.der_parse_strict(bytes.fromhex(sig), c)
        if rs is None:
            assert not flag, i
            n_der += 1
            continue
        r, s = rs
        if not (1 <= r < c.n and 1 <= s < c.n):
            assert not flag, i
            n_range += 1
            continue
        if c.low_s and s > c.n >> 1:
            # raw high-s must be rejected by the engine-level semantics ...
            assert not o.verify_prehash(c, Q, hf(bytes.fromhex(msg)).digest(), r, s)
            s = c.n - s      # ... the reference runner normalises first (k256/src/ecdsa.rs:389)
        n_math += 1
        assert o.verify_prehash(c, Q, hf(bytes.fromhex(msg)).digest(), r, s) == bool(flag), i
    assert n_der + n_range + n_math == len(blob["rows"])
    assert n_math > 130


def test_prehash_length_cases(golden):
    m = golden["misc"]["p256_prehash_sha384_verify"]          # longer than FB: keep leftmost FB
    assert o.verify_prehash(o.P256, (H(m["qx"]), H(m["qy"])), bytes.fromhex(m["prehash"]), H(m["r"]), H(m["s"]))
    m = golden["misc"]["p384_prehash_sha256_verify"]          # shorter than FB: left pad
    assert o.verify_prehash(o.P384, (H(m["qx"]), H(m["qy"])), bytes.fromhex(m["prehash"]), H(m["r"]), H(m["s"]))
    assert o.bits2field(o.P256, b"\x01" * 15) is None          # < FB/2 => Err


def test_p256_rfc6979(golden):
    m = golden["misc"]["p256_rfc6979"]
    Q = o.mul_gen(o.P256, H(m["d"]))
    for msg, sig in m["sigs"]:
        r, s = H(sig[:64]), H(sig[64:])
        assert o.verify_prehash(o.P256, Q, hashlib.sha256(msg.encode()).digest(), r, s)


def test_sec1(golden):
    for cname in ("k256", "p256"):
        c = o.curve(cname)
        unc, comp = golden["misc"][cname + "_sec1_generator"]["hexes"]
        assert o.sec1_encode(c, c.G, False).hex() == unc
        assert o.sec1_encode(c, c.G, True).hex() == comp
        for enc in (unc, comp):
            assert o.sec1_decode(c, bytes.fromhex(enc)) == (True, c.G)
        assert o.sec1_encode(c, None) == b"\x00"
        assert o.slot_encode(c, None, True) == b"\x00" * 33


def test_sm2(golden):
    m = golden["misc"]["sm2_pkcs8"]
    Q = o.mul_gen(o.SM2, H(m["d"]))
    assert o.sec1_encode(o.SM2, Q, False).hex() == m["sec1_public"]
    m = golden["misc"]["sm2dsa"]
    ok, Q = o.sec1_decode(o.SM2, bytes.fromhex(m["sec1_public"]))
    assert ok
    try:
        hashlib.new("sm3")
    except ValueError:
        pytest.skip("no sm3 in hashlib")
    z = o.sm2_z_hash(m["identity"].encode(), Q)
    e = hashlib.new("sm3", z + m["msg"].encode()).digest()
    r, s = H(m["sig"][:64]), H(m["sig"][64:])
    assert o.sm2dsa_verify_prehashed(Q, e, r, s)
    assert not o.sm2dsa_verify_prehashed(Q, e, r, s ^ 1)


def test_glv_and_recode():
    import random
    rng = random.Random(1)
    n = o.K256.n
    assert pow(o.K256_LAMBDA, 3, n) == 1 and pow(o.K256_BETA, 3, o.K256.p) == 1
    for k in [0, 1, 2, n - 1, n - 2, n >> 1, (n >> 1) + 1, 1 << 128, (1 << 128) - 1] + [rng.randrange(n) for _ in range(2000)]:
        a1, s1, a2, s2 = o.k256_decompose_signed(k)
        assert a1 < 1 << 128 and a2 < 1 << 128
        r1 = -a1 if s1 else a1
        r2 = -a2 if s2 else a2
        assert (r1 + r2 * o.K256_LAMBDA - k) % n == 0
        for a in (a1, a2):
            d = o.radix16_signed(a, 33)
            assert sum(x << (4 * i) for i, x in enumerate(d)) == a and all(-8 <= x <= 8 for x in d)


def test_batch_normalize_identity_slots():
    # k256 projective.rs:790-803,823-833: Z=0 slots -> IDENTITY, others unaffected
    c = o.K256
    P = o.mul_gen(c, 5)
    pts = [(P[0] * 7 % c.p, P[1] * 7 % c.p, 7), (0, 1, 0), (c.gx * 3 % c.p, c.gy * 3 % c.p, 3)]
    assert o.batch_normalize(c, pts) == [P, None, c.G]


# ---------------------------------------------------------------------------------------------
# "next" rows: signing, recovery, BIP340

@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "p192", "p224"])
def test_ecdsa_signing_vectors(golden, cname):
    # the FIPS / RFC vectors carry d and k: sign_prehashed must reproduce (r, s) (ecdsa_core new_signing_test!)
    c = o.curve(cname)
    for v in golden["ecdsa"][cname]["vectors"]:
        d, k, z = H(v["d"]), H(v["k"]), bytes.fromhex(v["m"])
        zb = o.bits2field(c, z)
        r, s, recid = o.sign_prehashed(c, d, k, zb)
        if c.low_s and H(v["s"]) > c.n >> 1:
            assert (r, c.n - s) == (H(v["r"]), H(v["s"]))
        else:
            assert (r, s) == (H(v["r"]), H(v["s"]))
        assert o.mul_gen(c, d) == (H(v["q_x"]), H(v["q_y"]))
        assert o.recover_from_prehash(c, z, r, s, recid) == (H(v["q_x"]), H(v["q_y"]))
        assert o.recover_from_prehash(c, z, r, s, recid ^ 1) != (H(v["q_x"]), H(v["q_y"]))


def test_p256_rfc6979_signing(golden):
    m = golden["misc"]["p256_rfc6979"]
    d = H(m["d"])
    for msg, sig in m["sigs"]:
        z = hashlib.sha256(msg.encode()).digest()
        k = o.rfc6979_k(o.P256, d, z)
        r, s, _ = o.sign_prehashed(o.P256, d, k, z)
        assert "%064x%064x" % (r, s) == sig


def test_k256_recovery_vectors(golden):
    c = o.K256
    for v in golden["next"]["k256_recovery"]["vectors"]:
        r, s = H(v["sig"][:64]), H(v["sig"][64:])
        Q = o.recover_from_prehash(c, bytes.fromhex(v["prehash"]), r, s, v["recid"])
        assert Q is not None and o.sec1_encode(c, Q, True).hex() == v["pk"]
        assert o.recover_from_prehash(c, bytes.fromhex(v["prehash"]), r, s, v["recid"] ^ 1) != Q
    # high-s signatures are rejected by the final verify step (k256/src/ecdsa.rs:203-205)
    v = golden["next"]["k256_recovery"]["vectors"][0]
    r, s = H(v["sig"][:64]), H(v["sig"][64:])
    assert o.recover_from_prehash(c, bytes.fromhex(v["prehash"]), r, c.n - s, v["recid"] ^ 1) is None
    # recid bit 1 with r + n >= p cannot decompress
    assert o.recover_from_prehash(c, bytes.fromhex(v["prehash"]), r, s, v["recid"] | 2) is None


def test_k256_ethereum_end_to_end(golden):
    c = o.K256
    e = golden["next"]["k256_ethereum_sign_recover"]
    d, z = H(e["d"]), bytes.fromhex(e["prehash"])
    k = o.rfc6979_k(c, d, z)
    r, s, recid = o.sign_prehashed(c, d, k, z)
    assert "%064x%064x" % (r, s) == e["sig"] and recid == e["recid"]
    assert o.recover_from_prehash(c, z, r, s, recid) == o.mul_gen(c, d)


def test_bip340_vectors(golden):
    for v in golden["next"]["bip340_sign"]["vectors"]:
        d = H(v["secret_key"])
        P = o.mul_gen(o.K256, d)
        assert "%064X" % P[0] == v["public_key"].upper()
        sig = o.schnorr_sign_prehash(d, bytes.fromhex(v["message"]), bytes.fromhex(v["aux_rand"]))
        assert sig.hex() == v["signature"].lower(), v["index"]
        assert o.schnorr_verify_prehash(bytes.fromhex(v["public_key"]), bytes.fromhex(v["message"]), sig)
    seen = set()
    for v in golden["next"]["bip340_verify"]["vectors"]:
        got = o.schnorr_verify_prehash(bytes.fromhex(v["public_key"]), bytes.fromhex(v["message"]), bytes.fromhex(v["signature"]))
        assert got == v["valid"], v["index"]
        seen.add(v["index"])
    assert seen == set(range(4, 15))


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "sm2", "p192", "p224"])
def test_decompress_roundtrip(cname):
    import random
    c = o.curve(cname)
    rng = random.Random(3)
    for _ in range(10):
        P = o.mul_gen(c, rng.randrange(1, c.n))
        assert o.decompress(c, P[0], P[1] & 1) == P
        assert o.decompress(c, P[0], (P[1] & 1) ^ 1) == (P[0], c.p - P[1])
    assert o.decompress(c, c.p, 0) is None
    bad = next(x for x in range(1, 50) if o.sqrt_mod(c, (x ** 3 + c.a * x + c.b) % c.p) is None)
    assert o.decompress(c, bad, 0) is None


def test_generated_constants_match_oracle():
    """tools/gen_consts.py carries its own copy of the curve parameters (the product's build tooling does not import
    oracle/): they must be the oracle's, and regenerating the header must reproduce the committed file."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_consts", os.path.join(root, "tools", "gen_consts.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    for name in ("K256", "P256", "P384", "SM2", "P192", "P224"):
        a, b = getattr(g.o, name), getattr(o, name)
        for f in ("p", "a", "b", "n", "gx", "gy", "fb", "cid", "compress", "low_s"):
            assert getattr(a, f) == getattr(b, f), (name, f)
    for f in ("K256_LAMBDA", "K256_BETA", "K256_MINUS_B1", "K256_MINUS_B2", "K256_G1", "K256_G2"):
        assert getattr(g.o, f) == getattr(o, f)


@pytest.mark.parametrize("cname", ["k256", "p256", "p384", "sm2", "p192", "p224"])
def test_decompact_conventions(cname):
    """k256 decompact = even root (k256 affine.rs:204-211); primeorder decompact = to_compact(decompress(x, 0)), the
    root with the smaller y (primeorder/src/affine.rs:66-77,148-156).  SEC1 tag 05 goes through the same function."""
    import random
    c = o.curve(cname)
    rng = random.Random(8 + c.cid)
    for _ in range(12):
        P = o.mul_gen(c, rng.randrange(1, c.n))
        D = o.decompact(c, P[0])
        assert D[0] == P[0] and D[1] in (P[1], c.p - P[1])
        if cname == "k256":
            assert D[1] % 2 == 0
        else:
            assert D[1] <= c.p - D[1]
        assert o.sec1_decode(c, b"\x05" + P[0].to_bytes(c.fb, "big")) == (True, D)
