"""Shared case builders for the SURVEY §8f rows (SEC1 decoding, recovery, BIP340, SM2DSA, signing): byte inputs plus
the oracle's expected outputs.  Used by the CPU logic tests (host emulation) and by the GPU parity tests."""
import hashlib
import random

from oracle import ecoracle as o


def be(v, n):
    return int(v).to_bytes(n, "big")


def non_residue_x(c):
    return next(x for x in range(1, 200) if o.sqrt_mod(c, (x ** 3 + c.a * x + c.b) % c.p) is None)


# ---------------------------------------------------------------------------------------------- decoding
def decode_cases(c, n_random=12, seed=5):
    """(slots, stride, expected_status, expected_xy) for DEC_SEC1 with stride 1+2FB (both encodings in one batch)."""
    fb = c.fb
    stride = 1 + 2 * fb
    rng = random.Random(seed + c.cid)
    slots, status, xy = [], [], []

    def add(enc, st, P):
        slots.append(enc + bytes(stride - len(enc)))
        status.append(st)
        xy.append(be(P[0], fb) + be(P[1], fb) if st == 1 else bytes(2 * fb))

    for i in range(n_random):
        P = o.mul_gen(c, rng.randrange(1, c.n))
        add(o.sec1_encode(c, P, True), 1, P)
        add(o.sec1_encode(c, P, False), 1, P)
        if i % 3 == 0:   # flipped parity tag decodes to the negated point
            add(bytes([5 - (2 + (P[1] & 1))]) + be(P[0], fb), 1, (P[0], c.p - P[1]))
    add(o.sec1_encode(c, c.G, True), 1, c.G)
    add(bytes(1), 2, None)                                                    # identity slot
    add(b"\x00" + bytes(fb - 1) + b"\x01", 0, None)                          # tag 0 with trailing garbage
    add(b"\x02" + be(c.p, fb), 0, None)                                      # x = p (non-canonical)
    add(b"\x03" + be((1 << (8 * fb)) - 1, fb), 0, None)                      # x = 2^(8FB) - 1
    add(b"\x02" + be(non_residue_x(c), fb), 0, None)                         # no square root
    add(b"\x04" + be(c.gx, fb) + be(c.gy ^ 1, fb), 0, None)                  # off curve
    add(b"\x04" + be(c.gx, fb) + be(c.p, fb), 0, None)                       # y = p
    for i in range(3):                                                        # tag 05 = SEC1 compact: decompact
        P = o.mul_gen(c, rng.randrange(1, c.n))
        add(b"\x05" + be(P[0], fb), 1, o.decompact(c, P[0]))
    add(b"\x05" + be(non_residue_x(c), fb), 0, None)
    add(b"\x06" + be(c.gx, fb), 0, None)                                     # unknown tags
    add(b"\x01" + be(c.gx, fb), 0, None)
    return b"".join(slots), stride, bytes(status), b"".join(xy)


def compact_cases(c, n_random=8, seed=6):
    fb = c.fb
    rng = random.Random(seed + c.cid)
    enc, status, xy = [], [], []
    for _ in range(n_random):
        P = o.mul_gen(c, rng.randrange(1, c.n))
        P = o.decompact(c, P[0])          # k256: even root; primeorder curves: the smaller y
        enc.append(be(P[0], fb)); status.append(1); xy.append(be(P[0], fb) + be(P[1], fb))
    for x in (c.p, non_residue_x(c)):
        enc.append(be(x, fb)); status.append(0); xy.append(bytes(2 * fb))
    return b"".join(enc), bytes(status), b"".join(xy)


# ---------------------------------------------------------------------------------------------- ECDSA helpers
def make_sig(c, rng, d=None):
    d = d or rng.randrange(1, c.n)
    z = be(rng.randrange(1 << (8 * c.fb)), c.fb)
    while True:
        k = rng.randrange(1, c.n)
        out = o.sign_prehashed(c, d, k, z)
        if out:
            return d, k, z, out


# ---------------------------------------------------------------------------------------------- recovery
def recover_cases(c, golden=None, n_random=10, seed=9):
    """rows (z, r, s, recid) -> expected SEC1 slot (curve default compression) and ok."""
    rng = random.Random(seed + c.cid)
    rows = []
    for i in range(n_random):
        d, k, z, (r, s, recid) = make_sig(c, rng)
        rows.append((z, r, s, recid))
        if i % 2 == 0:
            rows.append((z, r, s, recid ^ 1))          # wrong parity: another key (still "ok")
        if i % 3 == 0:
            rows.append((z, r, s, recid | 2))          # x-reduced bit without reduction: r + n >= p or no point
            rows.append((z, r, c.n - s, recid ^ 1))    # high-s twin: rejected on k256, fine elsewhere
        if i % 5 == 0:
            rows.append((z, 0, s, recid)); rows.append((z, r, 0, recid)); rows.append((z, c.n, s, recid)); rows.append((z, r, s, 4))
    from tests import crafted
    rows += crafted.reduced_x_recover_rows(c)          # x-reduced bit set AND recoverable (x(R) = r + n < p)
    # r whose restored x overflows / exceeds p, and an r with no point
    rows.append((bytes(c.fb), c.n - 1, 1, 2))
    rows.append((bytes(c.fb), non_residue_x(c), 1, 0))
    if golden is not None and c.name == "k256":
        for v in golden["next"]["k256_recovery"]["vectors"]:
            rows.append((bytes.fromhex(v["prehash"]), int(v["sig"][:64], 16), int(v["sig"][64:], 16), v["recid"]))
        e = golden["next"]["k256_ethereum_sign_recover"]
        rows.append((bytes.fromhex(e["prehash"]), int(e["sig"][:64], 16), int(e["sig"][64:], 16), e["recid"]))
    fb = c.fb
    lim = 1 << (8 * fb)
    zb = b"".join(r_[0] for r_ in rows)
    rsb = b"".join(be(r_[1] % lim, fb) + be(r_[2] % lim, fb) for r_ in rows)
    ids = bytes(r_[3] for r_ in rows)
    exp_keys, exp_ok = [], []
    for z, r, s, recid in rows:
        Q = o.recover_from_prehash(c, z, r, s, recid) if recid < 4 else None
        exp_ok.append(1 if Q is not None else 0)
        exp_keys.append(o.slot_encode(c, Q))
    return zb, rsb, ids, b"".join(exp_keys), bytes(exp_ok)


# ---------------------------------------------------------------------------------------------- BIP340
def schnorr_cases(golden, n_random=8, seed=12):
    """rows (pk_x, e_digest, sig) -> expected ok; all 15 BIP340 vectors + synthetic signatures and corruptions."""
    c = o.K256
    rows, exp = [], []

    def add(pk, msg, sig, e=None):
        e = e if e is not None else o.schnorr_challenge(sig[:32], pk, msg)
        rows.append((pk, e, sig))
        exp.append(1 if o.schnorr_verify_raw(pk, e, sig) else 0)

    for v in golden["next"]["bip340_sign"]["vectors"]:
        add(bytes.fromhex(v["public_key"]), bytes.fromhex(v["message"]), bytes.fromhex(v["signature"]))
    for v in golden["next"]["bip340_verify"]["vectors"]:
        add(bytes.fromhex(v["public_key"]), bytes.fromhex(v["message"]), bytes.fromhex(v["signature"]))
        assert exp[-1] == int(v["valid"])
    rng = random.Random(seed)
    for i in range(n_random):
        d = rng.randrange(1, c.n)
        msg = be(rng.getrandbits(256), 32)
        sig = o.schnorr_sign_prehash(d, msg, be(rng.getrandbits(256), 32))
        pk = be(o.mul_gen(c, d)[0], 32)
        add(pk, msg, sig)
        if i % 2 == 0:
            add(pk, msg, sig[:63] + bytes([sig[63] ^ 1]))
            add(pk, msg, bytes([sig[0] ^ 0x80]) + sig[1:])
            add(pk, be(int.from_bytes(msg, "big") ^ 1, 32), sig)
        if i % 4 == 0:
            add(pk, msg, sig[:32] + be(c.n, 32))                 # s = n
            add(pk, msg, sig[:32] + bytes(32))                   # s = 0
            add(pk, msg, be(c.p, 32) + sig[32:])                 # r = p
            add(pk, msg, bytes(32) + sig[32:])                   # r = 0
            add(be(non_residue_x(c), 32), msg, sig)              # key not on curve
            add(pk, msg, sig, e=bytes(32))                       # e = 0: R = s*G
            add(pk, msg, sig, e=be(c.n, 32))                     # e = n reduces to 0
            add(pk, msg, sig, e=be((1 << 256) - 1, 32))          # e >= n reduces once
    pkb = b"".join(r_[0] for r_ in rows)
    eb = b"".join(r_[1] for r_ in rows)
    sb = b"".join(r_[2] for r_ in rows)
    return pkb, eb, sb, bytes(exp)


# ---------------------------------------------------------------------------------------------- SM2DSA
def sm2dsa_sign(d, e_bytes, k):
    """SM2DSA signing (sm2/src/dsa/signing.rs) — only to manufacture valid rows for the verifier."""
    c = o.SM2
    e = o.reduce_once(c, int.from_bytes(e_bytes, "big"))
    x1 = o.mul_gen(c, k)[0]
    r = (e + x1) % c.n
    if r == 0 or r + k == c.n:
        return None
    s = pow(1 + d, -1, c.n) * (k - r * d) % c.n
    return (r, s) if s else None


def sm2dsa_cases(golden, n_random=10, seed=14):
    c = o.SM2
    rng = random.Random(seed)
    rows = []
    m = golden["misc"]["sm2dsa"]
    ok, Q = o.sec1_decode(c, bytes.fromhex(m["sec1_public"]))
    try:
        zA = o.sm2_z_hash(m["identity"].encode(), Q)
        e = hashlib.new("sm3", zA + m["msg"].encode()).digest()
        rows.append((Q, e, int(m["sig"][:64], 16), int(m["sig"][64:], 16)))
    except ValueError:
        pass
    for i in range(n_random):
        d = rng.randrange(1, c.n - 1)
        Q = o.mul_gen(c, d)
        e = be(rng.getrandbits(256), 32)
        sig = None
        while sig is None:
            sig = sm2dsa_sign(d, e, rng.randrange(1, c.n))
        r, s = sig
        rows.append((Q, e, r, s))
        if i % 2 == 0:
            rows.append((Q, e, r ^ 1, s)); rows.append((Q, e, r, s ^ 1)); rows.append((Q, be(int.from_bytes(e, "big") ^ 2, 32), r, s))
        if i % 4 == 0:
            rows.append((Q, e, 0, s)); rows.append((Q, e, r, 0)); rows.append((Q, e, c.n, s)); rows.append((Q, e, r, c.n - r))   # t = 0
            rows.append(((Q[0], Q[1] ^ 1), e, r, s))
    # identity lincomb: s*G + t*Q = O with Q = d*G needs s + t d = 0; then accept <=> r == e (mod n)
    d = rng.randrange(2, c.n - 1)
    Q = o.mul_gen(c, d)
    for _ in range(3):
        r = rng.randrange(1, c.n)
        s = (-r * d) * pow(1 + d, -1, c.n) % c.n          # s + (r + s) d = 0
        if s == 0 or (r + s) % c.n == 0:
            continue
        rows.append((Q, be(r, 32), r, s))                  # e == r: the reference accepts (x of IDENTITY is 0)
        rows.append((Q, be(r ^ 1, 32), r, s))
    qb = b"".join(be(r_[0][0], 32) + be(r_[0][1], 32) for r_ in rows)
    eb = b"".join(r_[1] for r_ in rows)
    lim = 1 << 256
    rsb = b"".join(be(r_[2] % lim, 32) + be(r_[3] % lim, 32) for r_ in rows)
    exp = bytes(1 if o.sm2dsa_verify_prehashed(r_[0], r_[1], r_[2], r_[3]) else 0 for r_ in rows)
    return qb, eb, rsb, exp


# ---------------------------------------------------------------------------------------------- signing
def sign_cases(c, golden=None, n_random=10, seed=17):
    rng = random.Random(seed + c.cid)
    rows = []
    if golden is not None and c.name in golden["ecdsa"]:
        for v in golden["ecdsa"][c.name]["vectors"]:
            rows.append((int(v["d"], 16), int(v["k"], 16), o.bits2field(c, bytes.fromhex(v["m"]))))
    if golden is not None and c.name == "k256":
        e = golden["next"]["k256_ethereum_sign_recover"]
        d, z = int(e["d"], 16), bytes.fromhex(e["prehash"])
        rows.append((d, o.rfc6979_k(c, d, z), z))
    for i in range(n_random):
        rows.append((rng.randrange(1, c.n), rng.randrange(1, c.n), be(rng.getrandbits(8 * c.fb), c.fb)))
    z = be(rng.getrandbits(8 * c.fb), c.fb)
    rows += [(0, 5, z), (5, 0, z), (c.n, 5, z), (5, c.n, z), (1, 1, bytes(c.fb)), (c.n - 1, c.n - 1, be((1 << (8 * c.fb)) - 1, c.fb)), (7, 1, z)]
    fb = c.fb
    lim = 1 << (8 * fb)
    db = b"".join(be(r_[0] % lim, fb) for r_ in rows)
    kb = b"".join(be(r_[1] % lim, fb) for r_ in rows)
    zb = b"".join(r_[2] for r_ in rows)
    rs, rid, ok = [], [], []
    for d, k, z_ in rows:
        out = o.sign_prehashed(c, d, k, z_) if 0 < d < c.n else None
        if out is None:
            rs.append(bytes(2 * fb)); rid.append(0); ok.append(0)
        else:
            rs.append(be(out[0], fb) + be(out[1], fb)); rid.append(out[2]); ok.append(1)
    return db, kb, zb, b"".join(rs), bytes(rid), bytes(ok)
