#!/usr/bin/env python3
"""Generate rustcrypto-elliptic-curves_b200/csrc/curve_consts.cuh (limb tables for every modulus and curve).

Constants come from SURVEY.md App. A / the reference (k256/src/arithmetic/mul.rs:129-152,
projective.rs:29-34, */src/arithmetic.rs, */src/lib.rs ORDER) and are embedded below; derived values (R, R^2, n0', b3 in
Montgomery form ...) are computed here with Python integers.  Usage: python tools/gen_consts.py
"""
import os
from types import SimpleNamespace as NS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Curve parameters (k256/src/arithmetic/field.rs:312-313, k256/src/lib.rs:76, k256 affine.rs:63-75,
# mul.rs:129-152, projective.rs:29-34; p256/src/arithmetic.rs:37-59, p256/src/lib.rs:74; p384/src/arithmetic.rs:36-61,
# p384/src/lib.rs:50; sm2/src/arithmetic.rs:37-58, sm2/src/lib.rs:60).  Self-contained on purpose: the product's
# build tooling does not import oracle/ (tests/test_oracle_golden.py::test_generated_constants_match_oracle compares).
_P256_P = 0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF
_P384_P = 2**384 - 2**128 - 2**96 + 2**32 - 1
_SM2_P = 0xFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF00000000FFFFFFFFFFFFFFFF
o = NS(
    K256=NS(name="k256", cid=0, p=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F, a=0, b=7,
            n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
            gx=0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
            gy=0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8, fb=32, compress=True, low_s=True),
    P256=NS(name="p256", cid=1, p=_P256_P, a=_P256_P - 3,
            b=0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B,
            n=0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551,
            gx=0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
            gy=0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5, fb=32, compress=False, low_s=False),
    P384=NS(name="p384", cid=2, p=_P384_P, a=_P384_P - 3,
            b=0xB3312FA7E23EE7E4988E056BE3F82D19181D9C6EFE8141120314088F5013875AC656398D8A2ED19D2A85C8EDD3EC2AEF,
            n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFC7634D81F4372DDF581A0DB248B0A77AECEC196ACCC52973,
            gx=0xAA87CA22BE8B05378EB1C71EF320AD746E1D3B628BA79B9859F741E082542A385502F25DBF55296C3A545E3872760AB7,
            gy=0x3617DE4A96262C6F5D9E98BF9292DC29F8F41DBD289A147CE9DA3113B5F0B8C00A60B1CE1D7E819D7A431D7C90EA0E5F, fb=48, compress=False, low_s=False),
    SM2=NS(name="sm2", cid=3, p=_SM2_P, a=_SM2_P - 3,
           b=0x28E9FA9E9D9F5E344D5A9E4BCF6509A7F39789F515AB8F92DDBCBD414D940E93,
           n=0xFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFF7203DF6B21C6052B53BBF40939D54123,
           gx=0x32C4AE2C1F1981195F9904466A39C9948FE30BBFF2660BE1715A4589334C74C7,
           gy=0xBC3736A2F4F6779C59BDCEE36B692153D0A9877CC62A474002DF32E52139F0A0, fb=32, compress=False, low_s=False),
    # p192/src/arithmetic.rs:36-55, p192/src/arithmetic/field.rs:43, p192/src/lib.rs:42 ("then p192 ... via the same template")
    P192=NS(name="p192", cid=4, p=2**192 - 2**64 - 1, a=2**192 - 2**64 - 1 - 3,
            b=0x64210519E59C80E70FA7E9AB72243049FEB8DEECC146B9B1,
            n=0xFFFFFFFFFFFFFFFFFFFFFFFF99DEF836146BC9B1B4D22831,
            gx=0x188DA80EB03090F67CBF20EB43A18800F4FF0AFD82FF1012,
            gy=0x07192B95FFC8DA78631011ED6B24CDD573F977A11E794811, fb=24, compress=False, low_s=False),
    # p224/src/arithmetic.rs (a = -3, b, generator), p224/src/arithmetic/field.rs (modulus, S = 96, generator 22), p224/src/lib.rs (order)
    P224=NS(name="p224", cid=5, p=2**224 - 2**96 + 1, a=2**224 - 2**96 + 1 - 3,
            b=0xB4050A850C04B3ABF54132565044B0B7D7BFD8BA270B39432355FFB4,
            n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFF16A2E0B8F03E13DD29455C5C2A3D,
            gx=0xB70E0CBD6BB4BF7F321390B94A03C1D356C21122343280D6115C1D21,
            gy=0xBD376388B5F723FB4C22DFE6CD4375A05A07476444D5819985007E34, fb=28, compress=False, low_s=False),
    K256_LAMBDA=0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72,
    K256_BETA=0x7AE96A2B657C07106E64479EAC3434E99CF0497512F58995C1396C28719501EE,
    K256_MINUS_B1=0xE4437ED6010E88286F547FA90ABFE4C3,
    K256_MINUS_B2=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFE8A280AC50774346DD765CDA83DB1562C,
    K256_G1=0x3086D221A7D46BCDE86C90E49284EB153DAA8A1471E8CA7FE893209A45DBB031,
    K256_G2=0xE4437ED6010E88286F547FA90ABFE4C4221208AC9DF506C61571B4AE8AC47F71,
)


def limbs(v, L):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(L)]


def arr(name, v, L):
    body = ", ".join("0x%08Xu" % x for x in limbs(v, L))
    return ("    ECB_HD static constexpr u32 %s(int i) { constexpr u32 t[%d] = {%s}; return t[i]; }\n"
            % (name, L, body))


def sparse_terms(m, L):
    """m = 2^(32L) + sum_k s_k 2^(32 e_k) - 1 with 0 < e_k < L, s_k = +-1, at most 4 middle terms; None otherwise."""
    import itertools
    d = m + 1 - (1 << (32 * L))
    for nt in range(1, 5):
        for es in itertools.combinations(range(1, L), nt):
            for ss in itertools.product((1, -1), repeat=nt):
                if sum(s << (32 * e) for s, e in zip(ss, es)) == d:
                    return list(zip(es, ss))
    return None


def nonresidue(m):
    g = 2
    while pow(g, (m - 1) // 2, m) != m - 1:
        g += 1
    return g


def mont_params(name, m, L, comment, base_field=False):
    R = 1 << (32 * L)
    n0 = (-pow(m, -1, 1 << 32)) % (1 << 32)
    s = "// %s\nstruct %s {\n    static constexpr int L = %d;\n    static constexpr u32 n0 = 0x%08Xu;\n" % (comment, name, L, n0)
    s += arr("p", m, L) + arr("one", R % m, L) + arr("r2", R * R % m, L)
    terms = sparse_terms(m, L) if n0 == 1 else None
    if terms:
        # ascending whole-width passes are valid when every pair of middle exponents sums to >= L (mont.cuh)
        asc = all(a[0] + b[0] >= L for i, a in enumerate(terms) for b in terms[i + 1:])
        desc = " ".join("%s2^%d" % ("+" if sg > 0 else "-", 32 * e) for e, sg in reversed(terms))
        s += "    // p = 2^%d %s - 1: Montgomery reduction by shifted additions (n0 = 1)\n" % (32 * L, desc)
        s += "    static constexpr bool SPARSE = true;\n    static constexpr int NT = %d;\n    static constexpr bool ASC_OK = %s;\n" % (
            len(terms), "true" if asc else "false")
        s += "    ECB_HD static constexpr int te(int k) { constexpr int t[%d] = {%s}; return t[k]; }\n" % (len(terms), ", ".join(str(e) for e, _ in terms))
        s += "    ECB_HD static constexpr int ts(int k) { constexpr int t[%d] = {%s}; return t[k]; }\n" % (len(terms), ", ".join(str(sg) for _, sg in terms))
        # -p folded into the first chain's high-half pass (mont.cuh redc_sparse): C = 2^(32(L+1)) - p = 1 (mod 2^(32 e0)); a '+'
        # chain adds C's limbs from e0 up, a '-' chain subtracts the negation of C >> 32 e0; the 1 at limb 0 is a carry-in
        e0, s0 = terms[0]
        C = (1 << (32 * (L + 1))) - m
        assert C % (1 << (32 * e0)) == 1
        words = L - e0 + 1
        V = C >> (32 * e0)
        Wv = V if s0 > 0 else (-V) % (1 << (32 * words))
        fold = [0] * e0 + [(Wv >> (32 * i)) & 0xFFFFFFFF for i in range(words)]
        s += "    // -p for the sign-based final correction: constants of the first chain's upper limbs and of the top word\n"
        s += "    ECB_HD static constexpr u32 foldc(int i) { constexpr u32 t[%d] = {%s}; return t[i]; }\n" % (L, ", ".join("0x%08Xu" % x for x in fold[:L]))
        s += "    static constexpr u32 foldc_top = 0x%08Xu;\n" % fold[L]
    else:
        s += "    static constexpr bool SPARSE = false;\n"
    # p = 2^(32L) - 2^(32e) + 1 (P-224): n0 = -1; reduction by one shifted addition chain and a subtraction (mont.cuh redc_negsparse)
    ne = next((e for e in range(1, L) if m == R - (1 << (32 * e)) + 1), None)
    if ne:
        s += "    // p = 2^%d - 2^%d + 1: Montgomery reduction without multiplications (n0 = -1)\n" % (32 * L, 32 * ne)
        s += "    static constexpr bool NEGSPARSE = true;\n    static constexpr int NE = %d;\n" % ne
    else:
        s += "    static constexpr bool NEGSPARSE = false;\n"
    if base_field and m % 4 == 1:
        # Tonelli-Shanks constants: m - 1 = 2^S * t, t odd; root = g^t in Montgomery form for the smallest non-residue g
        S, t = 0, m - 1
        while t % 2 == 0:
            S, t = S + 1, t // 2
        s += "    // p = 1 (mod 4): square roots by Tonelli-Shanks, p - 1 = 2^%d * t\n" % S
        s += "    static constexpr bool SQRT_TS = true;\n    static constexpr int TS_S = %d;\n" % S
        s += arr("ts_half_t", (t - 1) // 2, L) + arr("ts_root", pow(nonresidue(m), t, m) * R % m, L)
    else:
        s += "    static constexpr bool SQRT_TS = false;\n"
    s += "};\n\n"
    return s


def main():
    out = []
    out.append("// GENERATED by tools/gen_consts.py — do not edit.\n#pragma once\n#include \"bigint.cuh\"\n#include \"fp_k256.cuh\"\n#include \"mont.cuh\"\n\nnamespace ecb {\n\n")
    for c in (o.K256, o.P256, o.P384, o.SM2, o.P192, o.P224):
        L = c.fb // 4
        up = c.name.upper()
        if c.name != "k256":
            out.append(mont_params(up + "P", c.p, L, "%s base field modulus" % c.name, base_field=True))
        out.append(mont_params(up + "N", c.n, L, "%s group order" % c.name))
    # curve descriptors
    for c in (o.K256, o.P256, o.P384, o.SM2, o.P192, o.P224):
        L = c.fb // 4
        up = c.name.upper()
        R = 1 << (32 * L)
        s = "struct Curve%s {\n" % up
        s += "    static constexpr int ID = %d;\n    static constexpr int L = %d;\n    static constexpr int FB = %d;\n" % (c.cid, L, c.fb)
        s += "    static constexpr bool A_IS_ZERO = %s;\n" % ("true" if c.a == 0 else "false")
        s += "    static constexpr bool COMPRESS_DEFAULT = %s;\n    static constexpr bool LOW_S = %s;\n" % (
            "true" if c.compress else "false", "true" if c.low_s else "false")
        if c.name == "k256":
            s += "    typedef FpK256 F;\n"
            f = lambda v: v % c.p
        else:
            s += "    typedef Mont<%sP> F;\n" % up
            f = lambda v: v * R % c.p
        s += "    typedef Mont<%sN> Fn;\n" % up
        s += "    // field-domain constants (Montgomery form for the primeorder curves)\n"
        s += arr("gx", f(c.gx), L) + arr("gy", f(c.gy), L) + arr("b", f(c.b), L) + arr("b3", f(3 * c.b), L)
        s += "    // plain integers\n"
        s += arr("n", c.n, L) + arr("half_n", c.n >> 1, L) + arr("p", c.p, L) + arr("p_minus_n", c.p - c.n, L)
        if c.name == "k256":
            Rn = R % c.n
            s += "    static constexpr u32 B3_SMALL = 21u;\n"
            s += "    // GLV (k256/src/arithmetic/mul.rs:129-152, projective.rs:29-34); *_R = value * 2^256 mod n\n"
            s += arr("beta", o.K256_BETA, L)
            s += arr("g1", o.K256_G1, L) + arr("g2", o.K256_G2, L)
            s += arr("minus_b1_R", o.K256_MINUS_B1 * Rn % c.n, L)
            s += arr("minus_b2_R", o.K256_MINUS_B2 * Rn % c.n, L)
            s += arr("minus_lambda_R", (c.n - o.K256_LAMBDA) * Rn % c.n, L)
        s += "};\n\n"
        out.append(s)
    out.append("}  // namespace ecb\n")
    path = os.path.join(ROOT, "rustcrypto-elliptic-curves_b200", "csrc", "curve_consts.cuh")
    with open(path, "w") as fh:
        fh.write("".join(out))
    print("wrote", path)


if __name__ == "__main__":
    main()
