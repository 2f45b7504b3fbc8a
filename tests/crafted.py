"""Crafted ECDSA rows that drive the verifier through chosen (u1, u2) = (z/s, r/s): exceptional cases of the
Jacobian fast path (P + P, P + (-P), identity accumulators), GLV corner scalars, zero prehash."""
from oracle import ecoracle as o


def craft(c, u1, u2, d=1):
    """(Q, z_bytes, r, s) with Q = d*G such that the verifier computes u1*G + u2*Q.  r is x((u1 + u2 d)G) mod n when
    that point exists (valid signature unless the low-s rule bites), else r = 1 (must be rejected)."""
    n = c.n
    u1 %= n
    u2 %= n
    Q = o.mul_gen(c, d)
    R = o.mul_gen(c, (u1 + u2 * d) % n)
    r = (R[0] % n) if R is not None else 1
    if r == 0 or u2 == 0:
        return None
    s = r * pow(u2, -1, n) % n
    z = u1 * s % n
    return (Q, z.to_bytes(c.fb, "big"), r, s)


def exceptional_rows(c):
    n = c.n
    lam = o.K256_LAMBDA if c.name == "k256" else 3
    pairs = []
    for v in (1, 2, 3, 7, 8, 9, 15):
        pairs += [(v, v, 1), (v, n - v, 1), (v + (5 << 16), n - v, 1), (v, v, n - 1), (n - v, v, 1), (0, v, 1), (0, n - v, 1),
                  (v, 1, v), (v << 16, v << 16, 1), (v << 16, n - (v << 16), 1), ((v << 16) + 3, n - (v << 16), 1)]
    for u2 in (lam, lam + 1, n - lam, lam - 1, (lam * 8) % n, 1 << 127, (1 << 128) - 1, 1 << 128, n >> 1, (n >> 1) + 1):
        pairs += [(0, u2, 1), (5, u2, 7), (u2, u2, 1), (n - u2, u2, 1)]
    rows = []
    for u1, u2, d in pairs:
        row = craft(c, u1, u2, d)
        if row is not None:
            rows.append(row)
    return rows
