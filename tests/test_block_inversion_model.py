"""Model of the CTA-wide inversion of the Montgomery-trick kernels (csrc/kernels_impl.cuh, BlockInv): the shuffle /
shared-memory schedule restated over Python integers, checked against pow(a, -1, p).  The CUDA code itself is covered by
the GPU tier (batch_normalize / verify / sign parity); this keeps the schedule's algebra pinned on the CPU tier."""
import random

import pytest

from oracle import ecoracle as o

LANES, WARPS = 32, 4


def block_inverse(a, p):
    """a: one product per thread of a 128-thread CTA (non-zero mod p) -> the 128 inverses, following BlockInv::run."""
    assert len(a) == LANES * WARPS
    below, above, wprod = [0] * len(a), [0] * len(a), [0] * WARPS
    for w in range(WARPS):
        pre = a[w * LANES:(w + 1) * LANES]
        suf = list(pre)
        d = 1
        while d < LANES:      # Hillis-Steele; an out-of-range shuffle returns the lane's own value and is discarded by the select
            up = [pre[l - d] if l >= d else pre[l] for l in range(LANES)]
            pre = [pre[l] * up[l] % p if l >= d else pre[l] for l in range(LANES)]
            down = [suf[l + d] if l + d < LANES else suf[l] for l in range(LANES)]
            suf = [suf[l] * down[l] % p if l + d < LANES else suf[l] for l in range(LANES)]
            d <<= 1
        for l in range(LANES):
            below[w * LANES + l] = 1 if l == 0 else pre[l - 1]
            above[w * LANES + l] = 1 if l == LANES - 1 else suf[l + 1]
        wprod[w] = pre[LANES - 1]
    before, tot = [0] * WARPS, 1           # thread 0: Montgomery's trick over the warp products
    for k in range(WARPS):
        before[k] = tot
        tot = tot * wprod[k] % p
    inv = pow(tot, p - 2, p)               # the one inversion chain of the CTA
    winv = [0] * WARPS
    for k in range(WARPS - 1, -1, -1):
        winv[k] = inv * before[k] % p
        inv = inv * wprod[k] % p
    return [below[i] * above[i] % p * winv[i // LANES] % p for i in range(len(a))]


@pytest.mark.parametrize("cname", ["k256", "p256", "p384"])
@pytest.mark.parametrize("modulus", ["p", "n"])
def test_block_inverse_matches_fermat(cname, modulus):
    c = o.curve(cname)
    m = getattr(c, modulus)
    rng = random.Random(7)
    a = [rng.randrange(1, m) for _ in range(LANES * WARPS)]
    for i in (0, 31, 32, 95, 127):         # idle threads and Z = 0 slots enter the product as 1
        a[i] = 1
    a[64] = m - 1
    inv = block_inverse(a, m)
    assert all(x * y % m == 1 for x, y in zip(a, inv))
