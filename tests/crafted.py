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


def reduced_x_points(c, count=3):
    """Curve points whose affine x lies in [n, p): x mod n = x - n is tiny, so r = x - n takes the verifier's second
    candidate (r + n < p) and the recovery id's x-reduced bit.  Exists because p - n ~ 2^(bits/2) for these curves."""
    out, t = [], 0
    while len(out) < count:
        t += 1
        P = o.decompress(c, c.n + t, t & 1)
        if P is not None:
            out.append(P)
    return out


def reduced_x_rows(c):
    """Valid (Q, z, r, s) rows whose R = u1*G + u2*Q has x(R) >= n (no discrete log needed: Q = (R - u1*G)/u2),
    plus twins with r replaced by x(R) - n + 1 (must be rejected)."""
    import random
    rng = random.Random(123 + c.cid)
    rows = []
    for R in reduced_x_points(c):
        u1, u2 = rng.randrange(1, c.n), rng.randrange(1, c.n)
        T = o.pt_add(c, R, o.pt_neg(c, o.mul_gen(c, u1)))
        Q = o.pt_mul(c, pow(u2, -1, c.n), T)
        r = R[0] - c.n
        s = r * pow(u2, -1, c.n) % c.n
        z = u1 * s % c.n
        if c.low_s and s > c.n >> 1:
            s = c.n - s                       # (u1, u2) -> (-u1, -u2): R -> -R, same x, still valid for the same z
        rows.append((Q, z.to_bytes(c.fb, "big"), r, s))
        rows.append((Q, z.to_bytes(c.fb, "big"), r + 1, s))
    return rows


def reduced_x_recover_rows(c):
    """(z, r, s, recid) with the x-reduced bit set and a recoverable key: R has x(R) = r + n < p."""
    import random
    rng = random.Random(321 + c.cid)
    rows = []
    for R in reduced_x_points(c):
        r = R[0] - c.n
        s = rng.randrange(1, (c.n >> 1) + 1)
        z = rng.randrange(1 << (8 * c.fb)).to_bytes(c.fb, "big")
        rows.append((z, r, s, 2 | (R[1] & 1)))
        rows.append((z, r, s, R[1] & 1))      # same r without the bit: decompress(r) — a different point or none
    return rows


def fixed_base_digits(c, u1, gw):
    """Signed digits of u1 as jac.cuh add_fixed_base recodes it: gw-bit windows biased by 2^(gw-1) below the top one
    (digit in [-2^(gw-1), 2^(gw-1) - 1]), unsigned top window that absorbs the carry.  sum(d_w 2^(gw w)) == u1."""
    bits = 8 * c.fb
    nwin = (bits + gw - 1) // gw
    half = 1 << (gw - 1)
    kb = u1 + sum(half << (gw * w) for w in range(nwin - 1))
    digs = [((kb >> (gw * w)) & ((1 << gw) - 1)) - half for w in range(nwin - 1)] + [kb >> (gw * (nwin - 1))]
    assert sum(d << (gw * w) for w, d in enumerate(digs)) == u1
    return digs


def fixed_base_collision_rows(c, gw, windows=None, seed=77):
    """Rows whose accumulator MEETS the fixed-base table entry it is about to receive: the public-input kernels add the
    windows of u1*G (low to high) onto acc = u2*Q, so with Q = d*G and u2 = (+-d_j 2^(gw j) - sum_{w<j} d_w 2^(gw w)) / d the
    j-th gathered addition is P + P (sign +) or P + (-P) (sign -, leaving the identity for the windows above).  One pair of
    rows per window j of the gw-bit recoding; verdicts come from the oracle (valid signatures unless a rule bites)."""
    import random
    rng = random.Random(seed + c.cid + 1000 * gw)
    n = c.n
    rows = []
    nwin = (8 * c.fb + gw - 1) // gw
    for j in (windows if windows is not None else range(nwin)):
        for sign in (1, -1):
            for _ in range(50):
                u1, d = rng.randrange(1, n), rng.randrange(2, n)
                digs = fixed_base_digits(c, u1, gw)
                if digs[j] == 0:
                    continue
                partial = sum(dw << (gw * w) for w, dw in enumerate(digs[:j]))
                u2 = (sign * (digs[j] << (gw * j)) - partial) * pow(d, -1, n) % n
                row = craft(c, u1, u2, d) if u2 else None
                if row is not None and not (c.low_s and row[3] > n >> 1):      # a high s would be refused before any point arithmetic
                    rows.append(row)
                    break
    return rows
