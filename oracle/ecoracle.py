"""ecoracle — CPU restatement (Python big integers) of the reference's batched hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported, linked or executed by the
product (``rustcrypto-elliptic-curves_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` replays every golden vector the
reference holds for this path (ADD/MUL group vectors, field KATs incl. the risc0 8x32 KATs,
FIPS 186-4 ECDSA vectors with the flipped-s negative, all 1172 Wycheproof rows, the sm2 d->Q pair
and SM2DSA vector) from ``tests/golden/*.json`` (scraped from /root/reference by
``tests/golden/make_golden.py``).

Because every observable output of the path is canonical (SEC1 bytes of a normalised affine
point, or a boolean), the oracle works in plain affine big-integer arithmetic; functions cite
the reference file:line whose *behaviour* they restate (paths relative to /root/reference).
Third-party crates that are not vendored in the reference tree are restated from their published
behaviour: ecdsa 0.16.9 (hazmat::verify_prehashed, bits2field, Signature range checks),
elliptic-curve 0.13.8 (BatchInvert/BatchNormalize, sec1 EncodedPoint), crypto-bigint
0.5.5-risczero.0 — see SURVEY.md App. B.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

Point = Optional[Tuple[int, int]]  # None = identity


@dataclass(frozen=True)
class Curve:
    name: str
    cid: int          # C-ABI curve id (include/ecb200.h)
    p: int
    a: int
    b: int
    n: int
    gx: int
    gy: int
    fb: int           # field bytes
    compress: bool    # reference default for To EncodedPoint (k256/src/lib.rs:108-111 etc.)
    low_s: bool       # k256 rejects high-s (k256/src/ecdsa.rs:201-208)

    @property
    def G(self) -> Point:
        return (self.gx, self.gy)

    @property
    def bits(self) -> int:
        return self.fb * 8


# Constants: SURVEY.md App. A (k256/src/arithmetic/field.rs:312-313, k256/src/lib.rs:76,
# k256/src/arithmetic/affine.rs:63-75; p256/src/arithmetic.rs:37-59, p256/src/lib.rs:74;
# p384/src/arithmetic.rs:36-61, p384/src/lib.rs:50; sm2/src/arithmetic.rs:37-58, sm2/src/lib.rs:60)
K256 = Curve(
    "k256", 0,
    p=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F,
    a=0, b=7,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
    gx=0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
    gy=0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
    fb=32, compress=True, low_s=True)
P256 = Curve(
    "p256", 1,
    p=0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF,
    a=-3 % 0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF,
    b=0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B,
    n=0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551,
    gx=0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
    gy=0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5,
    fb=32, compress=False, low_s=False)
_P384_P = 2**384 - 2**128 - 2**96 + 2**32 - 1
P384 = Curve(
    "p384", 2,
    p=_P384_P, a=-3 % _P384_P,
    b=0xB3312FA7E23EE7E4988E056BE3F82D19181D9C6EFE8141120314088F5013875AC656398D8A2ED19D2A85C8EDD3EC2AEF,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFC7634D81F4372DDF581A0DB248B0A77AECEC196ACCC52973,
    gx=0xAA87CA22BE8B05378EB1C71EF320AD746E1D3B628BA79B9859F741E082542A385502F25DBF55296C3A545E3872760AB7,
    gy=0x3617DE4A96262C6F5D9E98BF9292DC29F8F41DBD289A147CE9DA3113B5F0B8C00A60B1CE1D7E819D7A431D7C90EA0E5F,
    fb=48, compress=False, low_s=False)
_SM2_P = 0xFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF00000000FFFFFFFFFFFFFFFF
SM2 = Curve(
    "sm2", 3,
    p=_SM2_P, a=-3 % _SM2_P,
    b=0x28E9FA9E9D9F5E344D5A9E4BCF6509A7F39789F515AB8F92DDBCBD414D940E93,
    n=0xFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFF7203DF6B21C6052B53BBF40939D54123,
    gx=0x32C4AE2C1F1981195F9904466A39C9948FE30BBFF2660BE1715A4589334C74C7,
    gy=0xBC3736A2F4F6779C59BDCEE36B692153D0A9877CC62A474002DF32E52139F0A0,
    fb=32, compress=False, low_s=False)

# p192/src/arithmetic.rs:36-55 (a = -3, b, generator), p192/src/arithmetic/field.rs:43 (modulus), p192/src/lib.rs:42 (order)
_P192_P = 2**192 - 2**64 - 1
P192 = Curve(
    "p192", 4,
    p=_P192_P, a=-3 % _P192_P,
    b=0x64210519E59C80E70FA7E9AB72243049FEB8DEECC146B9B1,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFF99DEF836146BC9B1B4D22831,
    gx=0x188DA80EB03090F67CBF20EB43A18800F4FF0AFD82FF1012,
    gy=0x07192B95FFC8DA78631011ED6B24CDD573F977A11E794811,
    fb=24, compress=False, low_s=False)

# p224/src/arithmetic.rs (a = -3, b, generator), p224/src/arithmetic/field.rs (modulus), p224/src/lib.rs (order)
_P224_P = 2**224 - 2**96 + 1
P224 = Curve(
    "p224", 5,
    p=_P224_P, a=-3 % _P224_P,
    b=0xB4050A850C04B3ABF54132565044B0B7D7BFD8BA270B39432355FFB4,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFF16A2E0B8F03E13DD29455C5C2A3D,
    gx=0xB70E0CBD6BB4BF7F321390B94A03C1D356C21122343280D6115C1D21,
    gy=0xBD376388B5F723FB4C22DFE6CD4375A05A07476444D5819985007E34,
    fb=28, compress=False, low_s=False)

CURVES = {c.name: c for c in (K256, P256, P384, SM2, P192, P224)}
BY_ID = {c.cid: c for c in CURVES.values()}


def curve(name_or_id) -> Curve:
    return CURVES[name_or_id] if isinstance(name_or_id, str) else BY_ID[name_or_id]


# ----------------------------------------------------------------------------------------------
# field helpers

def inv_mod(x: int, m: int) -> int:
    return pow(x, -1, m)


def sqrt_mod(c: Curve, v: int) -> Optional[int]:
    """sqrt for p = 3 (mod 4): k256 field.rs:220-255, p256 field.rs:385-411, p384 field.rs:95-117, sm2 field.rs `sqrt`,
    p192 field.rs:103-108; Tonelli-Shanks for p = 1 (mod 4) (P-224, p224/src/arithmetic/field.rs:103-233).  Either root
    may come back; callers select by parity / size."""
    p = c.p
    v %= p
    if p % 4 == 3:
        r = pow(v, (p + 1) // 4, p)
        return r if r * r % p == v else None
    if v == 0:
        return 0
    if pow(v, (p - 1) // 2, p) != 1:
        return None
    s, t = 0, p - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    g = 2
    while pow(g, (p - 1) // 2, p) != p - 1:
        g += 1
    z, x, b, m = pow(g, t, p), pow(v, (t + 1) // 2, p), pow(v, t, p), s
    while b != 1:
        k, t2 = 0, b
        while t2 != 1:
            t2, k = t2 * t2 % p, k + 1
        zz = pow(z, 1 << (m - k - 1), p)
        x, z = x * zz % p, zz * zz % p
        b, m = b * z % p, k
    return x


def on_curve(c: Curve, x: int, y: int) -> bool:
    return (y * y - (x * x * x + c.a * x + c.b)) % c.p == 0


# ----------------------------------------------------------------------------------------------
# group law (affine; identity = None).  Behavioural restatement of the complete formulas
# k256/src/arithmetic/projective.rs:96-274 and primeorder/src/point_arithmetic.rs:209-317:
# P+P == double(P), P+(-P) == identity, identity is neutral.

def pt_neg(c: Curve, P: Point) -> Point:
    return None if P is None else (P[0], (-P[1]) % c.p)


def pt_dbl(c: Curve, P: Point) -> Point:
    if P is None:
        return None
    x, y = P
    if y == 0:
        return None
    lam = (3 * x * x + c.a) * inv_mod(2 * y, c.p) % c.p
    x3 = (lam * lam - 2 * x) % c.p
    return (x3, (lam * (x - x3) - y) % c.p)


def pt_add(c: Curve, P: Point, Q: Point) -> Point:
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % c.p == 0:
            return None
        return pt_dbl(c, P)
    lam = (y2 - y1) * inv_mod(x2 - x1, c.p) % c.p
    x3 = (lam * lam - x1 - x2) % c.p
    return (x3, (lam * (x1 - x3) - y1) % c.p)


def _jac_dbl(c, X, Y, Z):
    p = c.p
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    YY = Y * Y % p
    S = 4 * X * YY % p
    M = (3 * X * X + c.a * pow(Z, 4, p)) % p
    X3 = (M * M - 2 * S) % p
    Y3 = (M * (S - X3) - 8 * YY * YY) % p
    Z3 = 2 * Y * Z % p
    return (X3, Y3, Z3)


def _jac_add_affine(c, X1, Y1, Z1, x2, y2):
    p = c.p
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % p
    U2 = x2 * Z1Z1 % p
    S2 = y2 * Z1 * Z1Z1 % p
    H = (U2 - X1) % p
    R = (S2 - Y1) % p
    if H == 0:
        if R == 0:
            return _jac_dbl(c, X1, Y1, Z1)
        return (1, 1, 0)
    HH = H * H % p
    HHH = H * HH % p
    V = X1 * HH % p
    X3 = (R * R - HHH - 2 * V) % p
    Y3 = (R * (V - X3) - Y1 * HHH) % p
    Z3 = Z1 * H % p
    return (X3, Y3, Z3)


def pt_mul(c: Curve, k: int, P: Point) -> Point:
    """k·P (k reduced mod n first, as a `Scalar` always is).  Result-equivalent to
    k256 mul.rs:342-393,443-445 and primeorder projective.rs:106-150 (outputs are canonical).
    Uses Jacobian double-and-add with a single inversion so that 2^16-element checks stay cheap."""
    k %= c.n
    if P is None or k == 0:
        return None
    x2, y2 = P
    X, Y, Z = 1, 1, 0
    for bit in bin(k)[2:]:
        X, Y, Z = _jac_dbl(c, X, Y, Z)
        if bit == "1":
            X, Y, Z = _jac_add_affine(c, X, Y, Z, x2, y2)
    if Z == 0:
        return None
    zi = inv_mod(Z, c.p)
    zi2 = zi * zi % c.p
    return (X * zi2 % c.p, Y * zi2 * zi % c.p)


def pt_lincomb(c: Curve, terms: Sequence[Tuple[Point, int]]) -> Point:
    """sum k_i·P_i — k256 mul.rs:313-393 (LinearCombinationExt), primeorder projective.rs:415-420."""
    acc = None
    for P, k in terms:
        acc = pt_add(c, acc, pt_mul(c, k, P))
    return acc


def mul_gen(c: Curve, k: int) -> Point:
    """MulByGenerator — k256 mul.rs:424-439; primeorder projective.rs:422-431."""
    return pt_mul(c, k, c.G)


def proj_to_affine(c: Curve, X: int, Y: int, Z: int) -> Point:
    """Homogeneous projective (x = X/Z, y = Y/Z) to affine; Z = 0 -> identity.
    k256 projective.rs:73-84; primeorder projective.rs:62-74."""
    if Z % c.p == 0:
        return None
    zi = inv_mod(Z, c.p)
    return (X * zi % c.p, Y * zi % c.p)


def batch_normalize(c: Curve, pts: Sequence[Tuple[int, int, int]]) -> List[Point]:
    """k256 projective.rs:350-379 / primeorder projective.rs:382-413 + BatchInvert (Montgomery trick,
    SURVEY App. B.6): zero Z replaced by ONE for the product chain, such slots become IDENTITY."""
    zs = [(z % c.p) or 1 for (_, _, z) in pts]
    pref = []
    acc = 1
    for z in zs:
        acc = acc * z % c.p
        pref.append(acc)
    out: List[Point] = [None] * len(pts)
    if not pts:
        return out
    inv = inv_mod(acc, c.p)
    for i in range(len(pts) - 1, -1, -1):
        zi = inv * (pref[i - 1] if i else 1) % c.p
        inv = inv * zs[i] % c.p
        X, Y, Z = pts[i]
        out[i] = None if Z % c.p == 0 else (X * zi % c.p, Y * zi % c.p)
    return out


# ----------------------------------------------------------------------------------------------
# SEC1 encoding (k256 affine.rs:233-238,272-284; primeorder affine.rs:233-243,340-358)

def sec1_encode(c: Curve, P: Point, compress: Optional[bool] = None) -> bytes:
    """EncodedPoint bytes: identity = single 00; compressed 02/03||x; uncompressed 04||x||y."""
    if compress is None:
        compress = c.compress
    if P is None:
        return b"\x00"
    x, y = P
    if compress:
        return bytes([2 + (y & 1)]) + x.to_bytes(c.fb, "big")
    return b"\x04" + x.to_bytes(c.fb, "big") + y.to_bytes(c.fb, "big")


def slot_encode(c: Curve, P: Point, compress: Optional[bool] = None) -> bytes:
    """Fixed-stride output slot used by the C ABI (include/ecb200.h): SEC1 bytes zero-padded to
    1+FB (compressed) or 1+2FB (uncompressed); identity = all zeros (== GroupEncoding::to_bytes
    identity, k256 affine.rs:233-238)."""
    if compress is None:
        compress = c.compress
    size = 1 + (c.fb if compress else 2 * c.fb)
    e = sec1_encode(c, P, compress)
    return e + b"\x00" * (size - len(e))


def sec1_decode(c: Curve, data: bytes) -> Tuple[bool, Point]:
    """from_encoded_point (k256 affine.rs:241-270; primeorder affine.rs:164-195) incl. decompress
    (k256 affine.rs:184-202; primeorder affine.rs:129-150).  Returns (ok, point)."""
    if len(data) == 1 and data[0] == 0:
        return True, None
    if len(data) == 1 + c.fb and data[0] in (2, 3):
        x = int.from_bytes(data[1:], "big")
        if x >= c.p:
            return False, None
        y = sqrt_mod(c, (x * x * x + c.a * x + c.b) % c.p)
        if y is None:
            return False, None
        if (y & 1) != (data[0] & 1):
            y = c.p - y
        return True, (x, y)
    if len(data) == 1 + c.fb and data[0] == 5:            # sec1::Tag::Compact -> decompact (x only)
        P = decompact(c, int.from_bytes(data[1:], "big"))
        return (P is not None), P
    if len(data) == 1 + 2 * c.fb and data[0] == 4:
        x = int.from_bytes(data[1:1 + c.fb], "big")
        y = int.from_bytes(data[1 + c.fb:], "big")
        if x >= c.p or y >= c.p or not on_curve(c, x, y):
            return False, None
        return True, (x, y)
    return False, None


# ----------------------------------------------------------------------------------------------
# k256 GLV + signed radix-16 (restated so the device decomposition can be checked digit-for-digit)

K256_LAMBDA = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72  # mul.rs:4-5
K256_BETA = 0x7AE96A2B657C07106E64479EAC3434E99CF0497512F58995C1396C28719501EE    # projective.rs:29-34
K256_MINUS_B1 = 0xE4437ED6010E88286F547FA90ABFE4C3                                  # mul.rs:135-138
K256_MINUS_B2 = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFE8A280AC50774346DD765CDA83DB1562C  # mul.rs:140-143
K256_G1 = 0x3086D221A7D46BCDE86C90E49284EB153DAA8A1471E8CA7FE893209A45DBB031        # mul.rs:145-148
K256_G2 = 0xE4437ED6010E88286F547FA90ABFE4C4221208AC9DF506C61571B4AE8AC47F71        # mul.rs:150-152


def k256_mul_shift_384(a: int, b: int) -> int:
    """WideScalar::mul_shift_vartime(384) — wide64.rs:64-119: (a*b) >> 384, plus the bit below."""
    prod = a * b
    return (prod >> 384) + ((prod >> 383) & 1)


def k256_decompose(k: int) -> Tuple[int, int]:
    """decompose_scalar — mul.rs:260-268.  Returns (r1, r2) as scalars mod n with
    r1 + r2*lambda == k (mod n)."""
    n = K256.n
    c1 = k256_mul_shift_384(k, K256_G1) * K256_MINUS_B1 % n
    c2 = k256_mul_shift_384(k, K256_G2) * K256_MINUS_B2 % n
    r2 = (c1 + c2) % n
    r1 = (k + r2 * (n - K256_LAMBDA)) % n
    return r1, r2


def k256_decompose_signed(k: int) -> Tuple[int, bool, int, bool]:
    """mul.rs:350-362: (|r1|, r1 negated?, |r2|, r2 negated?) with |ri| < 2^128."""
    n = K256.n
    r1, r2 = k256_decompose(k)
    s1 = r1 > n >> 1
    s2 = r2 > n >> 1
    return (n - r1 if s1 else r1, s1, n - r2 if s2 else r2, s2)


def radix16_signed(v: int, digits: int) -> List[int]:
    """Radix16Decomposition::new — mul.rs:281-304 (digits in -8..7, last possibly 8)."""
    d = [(v >> (4 * i)) & 15 for i in range(digits - 1)] + [0]
    # reference fills D-1 nibbles (2 per byte) and leaves the last digit for the carry
    for i in range(digits - 1):
        carry = (d[i] + 8) >> 4
        d[i] -= carry << 4
        d[i + 1] += carry
    return d


def k256_lincomb_reference_algorithm(terms: Sequence[Tuple[Point, int]]) -> Point:
    """Step-for-step restatement of lincomb() — mul.rs:342-393 (tables, digits, 33 rounds), used to
    prove that the restated GLV constants / recoding reproduce plain k·P."""
    c = K256
    tables = []
    digs = []
    for P, k in terms:
        a1, s1, a2, s2 = k256_decompose_signed(k % c.n)
        assert a1 < 1 << 128 and a2 < 1 << 128
        P1 = pt_neg(c, P) if s1 else P
        Pb = None if P is None else (P[0] * K256_BETA % c.p, P[1])
        P2 = pt_neg(c, Pb) if s2 else Pb
        for Pt, a in ((P1, a1), (P2, a2)):
            t = [None]
            for j in range(8):
                t.append(pt_add(c, t[-1], Pt))
            tables.append(t)
            digs.append(radix16_signed(a, 33))
    acc = None
    for i in range(32, -1, -1):
        for _ in range(4):
            acc = pt_dbl(c, acc)
        for t, d in zip(tables, digs):
            e = t[abs(d[i])]
            acc = pt_add(c, acc, pt_neg(c, e) if d[i] < 0 else e)
    return acc


# ----------------------------------------------------------------------------------------------
# ECDSA verify (ecdsa 0.16.9 hazmat::verify_prehashed + bits2field,
# call sites k256/src/ecdsa.rs:200-209, p256/src/ecdsa.rs:71-75)

def bits2field(c: Curve, prehash: bytes) -> Optional[bytes]:
    """len < FB/2 -> Err; shorter -> left-pad with zeros; longer -> keep leftmost FB bytes."""
    if len(prehash) < c.fb // 2:
        return None
    if len(prehash) < c.fb:
        return b"\x00" * (c.fb - len(prehash)) + prehash
    return prehash[:c.fb]


def reduce_once(c: Curve, v: int) -> int:
    """Reduce<Uint>::reduce — k256 scalar.rs:700-713, p256 scalar.rs:661-673: one conditional
    subtraction of n (valid because 2^(8FB) < 2n)."""
    return v - c.n if v >= c.n else v


def verify_prehashed(c: Curve, Q: Point, z_bytes: bytes, r: int, s: int) -> bool:
    """verify_prehashed(&Q, z, sig): z_bytes is the FB-byte output of bits2field.  Range failures of
    r/s (rejected at Signature construction in the reference) and invalid public keys (rejected at
    VerifyingKey construction) return False — in the batch ABI those are per-element data."""
    if not (1 <= r < c.n and 1 <= s < c.n):
        return False
    if Q is None or not (0 <= Q[0] < c.p and 0 <= Q[1] < c.p) or not on_curve(c, Q[0], Q[1]):
        return False
    if c.low_s and s > c.n >> 1:           # k256/src/ecdsa.rs:203-205, scalar.rs:519-523
        return False
    z = reduce_once(c, int.from_bytes(z_bytes, "big"))
    w = inv_mod(s, c.n)
    u1 = z * w % c.n
    u2 = r * w % c.n
    R = pt_lincomb(c, [(c.G, u1), (Q, u2)])
    x = 0 if R is None else R[0]           # AffinePoint::IDENTITY.x == 0 (k256 affine.rs:51-55)
    return reduce_once(c, x) == r


def verify_prehash(c: Curve, Q: Point, prehash: bytes, r: int, s: int) -> bool:
    """VerifyingKey::verify_prehash (PrehashVerifier) = bits2field + verify_prehashed."""
    z = bits2field(c, prehash)
    return False if z is None else verify_prehashed(c, Q, z, r, s)


def sm2dsa_verify_prehashed(Q: Point, e_bytes: bytes, r: int, s: int) -> bool:
    """SM2DSA verify — sm2/src/dsa/verifying.rs:130-168 (prehash must be exactly 32 bytes)."""
    c = SM2
    if len(e_bytes) != 32:
        return False
    if not (1 <= r < c.n and 1 <= s < c.n):
        return False
    e = reduce_once(c, int.from_bytes(e_bytes, "big"))
    t = (r + s) % c.n
    if t == 0:
        return False
    R = pt_lincomb(c, [(c.G, s), (Q, t)])
    x1 = 0 if R is None else R[0]
    return r == (e + reduce_once(c, x1)) % c.n


def sm2_z_hash(distid: bytes, Q: Point) -> bytes:
    """Z = SM3(ENTL || ID || a || b || xG || yG || xA || yA) — sm2/src/lib.rs / distid hashing."""
    c = SM2
    h = hashlib.new("sm3")
    h.update((len(distid) * 8).to_bytes(2, "big") + distid)
    for v in (c.a, c.b, c.gx, c.gy, Q[0], Q[1]):
        h.update(v.to_bytes(32, "big"))
    return h.digest()


# ----------------------------------------------------------------------------------------------
# "next" rows of the hot path: point decompression, public-key recovery, BIP340 Schnorr
# verification, signing.  Third-party bodies (ecdsa 0.16.9 recovery.rs / hazmat.rs, rfc6979 0.4) are restated
# and anchored on the reference's own vectors (k256/src/ecdsa.rs:278-343, k256/src/schnorr.rs:217-445,
# */src/test_vectors/ecdsa.rs, p256/src/ecdsa.rs:98-118).

def decompress(c: Curve, x: int, y_is_odd: int) -> Optional[Point]:
    """DecompressPoint::decompress — k256 affine.rs:184-202, primeorder affine.rs:129-150: x >= p or a
    non-residue right-hand side -> None; otherwise the root whose parity matches."""
    if not 0 <= x < c.p:
        return None
    y = sqrt_mod(c, (x * x * x + c.a * x + c.b) % c.p)
    if y is None:
        return None
    if (y & 1) != (y_is_odd & 1):
        y = (c.p - y) % c.p
    return (x, y)


def decompact(c: Curve, x: int) -> Optional[Point]:
    """DecompactPoint::decompact.  k256 (affine.rs:204-211) takes the EVEN root (Taproot / BIP340 convention); the
    primeorder curves (primeorder/src/affine.rs:148-156 + to_compact :66-77) take the root with the SMALLER integer y."""
    P = decompress(c, x, 0)
    if P is None or c.name == "k256":
        return P
    return (P[0], min(P[1], c.p - P[1]))


def sign_prehashed(c: Curve, d: int, k: int, z_bytes: bytes) -> Optional[Tuple[int, int, int]]:
    """hazmat::sign_prehashed (ecdsa 0.16.9) + the k256 wrapper (k256/src/ecdsa.rs:181-198):
    R = k*G, r = x(R) mod n, s = k^-1 (z + r d); recid = y_odd | x_reduced << 1; k256 normalises s to the low
    half and flips the parity bit when it does.  None <=> Err (k = 0, r = 0 or s = 0).  d, k in [0, n)."""
    if not (0 < k < c.n):
        return None
    z = reduce_once(c, int.from_bytes(z_bytes, "big"))
    R = mul_gen(c, k)
    r = reduce_once(c, R[0])
    x_reduced = 1 if r != R[0] else 0
    s = inv_mod(k, c.n) * (z + r * d) % c.n
    if r == 0 or s == 0:
        return None
    y_odd = R[1] & 1
    if c.low_s and s > c.n >> 1:
        s = c.n - s
        y_odd ^= 1
    return r, s, y_odd | (x_reduced << 1)


def rfc6979_k(c: Curve, d: int, z_bytes: bytes, hashname: str = "sha256", ad: bytes = b"") -> int:
    """rfc6979::generate_k as driven by SignPrimitive::try_sign_prehashed_rfc6979: HMAC-DRBG over
    int2octets(d) || bits2octets(z) || ad, candidates until 0 < k < n.  Host-side in the reference too."""
    import hmac
    hl = hashlib.new(hashname).digest_size
    x = d.to_bytes(c.fb, "big")
    h = reduce_once(c, int.from_bytes(z_bytes, "big")).to_bytes(c.fb, "big")
    V, K = b"\x01" * hl, b"\x00" * hl
    mac = lambda key, msg: hmac.new(key, msg, hashname).digest()
    K = mac(K, V + b"\x00" + x + h + ad); V = mac(K, V)
    K = mac(K, V + b"\x01" + x + h + ad); V = mac(K, V)
    while True:
        T = b""
        while len(T) < c.fb:
            V = mac(K, V)
            T += V
        k = int.from_bytes(T[:c.fb], "big")
        if 0 < k < c.n:
            return k
        K = mac(K, V + b"\x00"); V = mac(K, V)


def recover_from_prehash(c: Curve, prehash: bytes, r: int, s: int, recid: int) -> Optional[Point]:
    """VerifyingKey::recover_from_prehash (ecdsa 0.16.9 recovery.rs; call sites k256/src/ecdsa.rs:300-340).
    recid bit 0 = y of R is odd, bit 1 = x of R was reduced mod n.  None <=> Err."""
    if not (1 <= r < c.n and 1 <= s < c.n) or not 0 <= recid <= 3:
        return None
    zb = bits2field(c, prehash)
    if zb is None:
        return None
    z = reduce_once(c, int.from_bytes(zb, "big"))
    x = r
    if recid & 2:
        x = r + c.n
        if x >> (8 * c.fb):          # checked_add overflow
            return None
    R = decompress(c, x, recid & 1)  # x >= p fails inside (FieldElement::from_bytes)
    if R is None:
        return None
    r_inv = inv_mod(r, c.n)
    u1 = (-(r_inv * z)) % c.n
    u2 = r_inv * s % c.n
    Q = pt_lincomb(c, [(c.G, u1), (R, u2)])
    if Q is None:                    # VerifyingKey::from_affine rejects the identity
        return None
    if not verify_prehashed(c, Q, zb, r, s):   # "ensure signature verifies with the recovered key"
        return None
    return Q


def tagged_hash(tag: bytes, *parts: bytes) -> bytes:
    """BIP340 tagged hash (k256/src/schnorr.rs:180-186)."""
    th = hashlib.sha256(tag).digest()
    h = hashlib.sha256(th + th)
    for p_ in parts:
        h.update(p_)
    return h.digest()


def schnorr_challenge(r_bytes: bytes, pk_bytes: bytes, msg: bytes) -> bytes:
    """The 32-byte digest whose reduction is e (k256/src/schnorr/verifying.rs:69-75); hashing stays on the host."""
    return tagged_hash(b"BIP0340/challenge", r_bytes, pk_bytes, msg)


def schnorr_verify_raw(pk_x: bytes, e_bytes: bytes, sig: bytes) -> bool:
    """BIP340 verification after hashing — k256/src/schnorr/verifying.rs:35-45 (key = lift_x, even y),
    schnorr.rs:143-160 (signature parsing: r < p, r != 0, s in [1, n-1]), verifying.rs:63-89:
    R = s*G - e*P; accept <=> R != identity, y(R) even, x(R) == r."""
    c = K256
    if len(pk_x) != 32 or len(sig) != 64 or len(e_bytes) != 32:
        return False
    P = decompress(c, int.from_bytes(pk_x, "big"), 0)
    if P is None:
        return False
    r, s = int.from_bytes(sig[:32], "big"), int.from_bytes(sig[32:], "big")
    if not (0 < r < c.p) or not (1 <= s < c.n):
        return False
    e = reduce_once(c, int.from_bytes(e_bytes, "big"))
    R = pt_lincomb(c, [(c.G, s), (P, (c.n - e) % c.n)])
    return R is not None and (R[1] & 1) == 0 and R[0] == r


def schnorr_verify_prehash(pk_x: bytes, msg32: bytes, sig: bytes) -> bool:
    if len(msg32) != 32 or len(sig) != 64:
        return False
    return schnorr_verify_raw(pk_x, schnorr_challenge(sig[:32], pk_x, msg32), sig)


def schnorr_sign_prehash(d: int, msg32: bytes, aux: bytes) -> bytes:
    """BIP340 signing with auxiliary randomness (k256/src/schnorr/signing.rs) — used to replay the signing vectors."""
    c = K256
    P = mul_gen(c, d)
    if P[1] & 1:
        d = c.n - d
    t = (d ^ int.from_bytes(tagged_hash(b"BIP0340/aux", aux), "big")).to_bytes(32, "big")
    pk = P[0].to_bytes(32, "big")
    k0 = int.from_bytes(tagged_hash(b"BIP0340/nonce", t, pk, msg32), "big") % c.n
    R = mul_gen(c, k0)
    k = c.n - k0 if R[1] & 1 else k0
    rb = R[0].to_bytes(32, "big")
    e = int.from_bytes(schnorr_challenge(rb, pk, msg32), "big") % c.n
    return rb + ((k + e * d) % c.n).to_bytes(32, "big")


# ----------------------------------------------------------------------------------------------
# strict DER (ecdsa::der::Signature::from_der as exercised by the Wycheproof runners)

def der_parse_strict(sig: bytes, c: Curve) -> Optional[Tuple[int, int]]:
    """SEQUENCE{INTEGER r, INTEGER s}, minimal definite lengths, minimal non-negative INTEGERs, no
    trailing bytes; values longer than FB bytes are rejected.  Returns (r, s) or None."""
    def read_len(buf, i):
        if i >= len(buf):
            return None
        b0 = buf[i]
        i += 1
        if b0 < 0x80:
            return b0, i
        nb = b0 & 0x7F
        if nb == 0 or nb > 4 or i + nb > len(buf):
            return None
        v = int.from_bytes(buf[i:i + nb], "big")
        if buf[i] == 0 or v < 0x80:
            return None                      # non-minimal length
        return v, i + nb

    def read_int(buf, i):
        if i >= len(buf) or buf[i] != 0x02:
            return None
        r = read_len(buf, i + 1)
        if r is None:
            return None
        ln, i = r
        if ln == 0 or i + ln > len(buf):
            return None
        body = buf[i:i + ln]
        if body[0] & 0x80:
            return None                      # negative
        if ln > 1 and body[0] == 0 and not (body[1] & 0x80):
            return None                      # non-minimal
        v = int.from_bytes(body, "big")
        if len(body.lstrip(b"\x00")) > c.fb:
            return None
        return v, i + ln

    if len(sig) < 2 or sig[0] != 0x30:
        return None
    r = read_len(sig, 1)
    if r is None:
        return None
    ln, i = r
    if i + ln != len(sig):
        return None
    a = read_int(sig, i)
    if a is None:
        return None
    rv, i = a
    b = read_int(sig, i)
    if b is None:
        return None
    sv, i = b
    if i != len(sig):
        return None
    return rv, sv


# ----------------------------------------------------------------------------------------------
# batch helpers on ABI byte layouts (what tests compare the CUDA path with)

def be(v: int, n: int) -> bytes:
    return v.to_bytes(n, "big")


def batch_mul_gen(c: Curve, k_bytes: bytes, compress=None) -> bytes:
    fb = c.fb
    out = bytearray()
    for i in range(len(k_bytes) // fb):
        k = int.from_bytes(k_bytes[i * fb:(i + 1) * fb], "big")
        out += slot_encode(c, mul_gen(c, k), compress)
    return bytes(out)


def batch_mul_var_affine(c: Curve, pts: bytes, inf: Optional[bytes], k_bytes: bytes, compress=None) -> bytes:
    fb = c.fb
    n = len(k_bytes) // fb
    out = bytearray()
    for i in range(n):
        x = int.from_bytes(pts[2 * fb * i:2 * fb * i + fb], "big")
        y = int.from_bytes(pts[2 * fb * i + fb:2 * fb * (i + 1)], "big")
        P = None if (inf is not None and inf[i]) else (x, y)
        k = int.from_bytes(k_bytes[i * fb:(i + 1) * fb], "big")
        out += slot_encode(c, pt_mul(c, k, P), compress)
    return bytes(out)


def batch_mul_var_proj(c: Curve, xyz: bytes, k_bytes: bytes, compress=None) -> bytes:
    fb = c.fb
    n = len(k_bytes) // fb
    out = bytearray()
    for i in range(n):
        X, Y, Z = (int.from_bytes(xyz[(3 * i + j) * fb:(3 * i + j + 1) * fb], "big") for j in range(3))
        P = proj_to_affine(c, X, Y, Z)
        k = int.from_bytes(k_bytes[i * fb:(i + 1) * fb], "big")
        out += slot_encode(c, pt_mul(c, k, P), compress)
    return bytes(out)


def batch_verify(c: Curve, q: bytes, z: bytes, rs: bytes) -> bytes:
    fb = c.fb
    n = len(z) // fb
    out = bytearray()
    for i in range(n):
        qx = int.from_bytes(q[2 * fb * i:2 * fb * i + fb], "big")
        qy = int.from_bytes(q[2 * fb * i + fb:2 * fb * (i + 1)], "big")
        r = int.from_bytes(rs[2 * fb * i:2 * fb * i + fb], "big")
        s = int.from_bytes(rs[2 * fb * i + fb:2 * fb * (i + 1)], "big")
        out.append(1 if verify_prehashed(c, (qx, qy), z[i * fb:(i + 1) * fb], r, s) else 0)
    return bytes(out)
