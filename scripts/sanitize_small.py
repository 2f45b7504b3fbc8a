"""Small end-to-end pass over every entry point at a batch size that is not a multiple of the CTA size (tail guards),
checked against the oracle.  Written for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`; that tool is
closed on this GPU pool, so in round 1 it only ran plain (passes)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecb200
from oracle import ecoracle as o
from tests import nextrows
from tests.conftest import load_golden

golden = {n: load_golden(n) for n in ("ecdsa", "misc", "next")}
eng = ecb200.Engine(0)
rng = random.Random(1)
for cname in ("k256", "p256", "p384", "sm2"):
    c = o.curve(cname)
    fb = c.fb
    n = 133      # not a multiple of the CTA size: exercises the tail guards
    ks = b"".join(rng.randrange(c.n).to_bytes(fb, "big") for _ in range(n))
    a = eng.mul_by_generator_batch(cname, ks, ecb200.FLAG_CT | ecb200.FLAG_UNCOMPRESSED)
    b = eng.mul_by_generator_batch(cname, ks, ecb200.FLAG_UNCOMPRESSED)
    assert a == b == o.batch_mul_gen(c, ks, False)
    pts = b"".join(a[(1 + 2 * fb) * i + 1:(1 + 2 * fb) * (i + 1)] for i in range(n))
    for fl in (ecb200.FLAG_CT, 0):
        out, inv = eng.mul_batch(cname, pts, ks, None, fl)
        assert out == o.batch_mul_var_affine(c, pts, None, ks) and not any(inv)
    xy, inf = eng.batch_normalize(cname, b"".join((pts[2 * fb * i:2 * fb * (i + 1)] + (1).to_bytes(fb, "big")) for i in range(n)))
    assert xy == pts and not any(inf)
    assert eng.lincomb(cname, pts, ks) == o.slot_encode(c, o.pt_lincomb(c, [((int.from_bytes(pts[2 * fb * i:2 * fb * i + fb], "big"),
                                                                               int.from_bytes(pts[2 * fb * i + fb:2 * fb * (i + 1)], "big")),
                                                                              int.from_bytes(ks[fb * i:fb * (i + 1)], "big")) for i in range(n)]))
    slots, stride, st, xy2 = nextrows.decode_cases(c, n_random=6)
    assert eng.decode_points(cname, slots, stride, 0) == (xy2, st)
    zb, rsb, ids, ekeys, eok = nextrows.recover_cases(c, golden if cname == "k256" else None, n_random=6)
    assert eng.ecdsa_recover(cname, zb, rsb, ids) == (ekeys, eok)
    db, kb, zb2, rs, rid, okx = nextrows.sign_cases(c, None, n_random=6)
    assert eng.ecdsa_sign(cname, db, kb, zb2) == (rs, rid, okx)
    good = [i for i in range(len(okx)) if okx[i]]
    q = b"".join(v.to_bytes(fb, "big") for i in good for v in o.mul_gen(c, int.from_bytes(db[fb * i:fb * i + fb], "big")))
    assert eng.ecdsa_verify(cname, q, b"".join(zb2[fb * i:fb * i + fb] for i in good), b"".join(rs[2 * fb * i:2 * fb * (i + 1)] for i in good)) == b"\x01" * len(good)
pk, e, sg, exp = nextrows.schnorr_cases(golden, n_random=4)
assert eng.schnorr_verify(pk, e, sg) == exp
q, e, rs, exp = nextrows.sm2dsa_cases(golden, n_random=4)
assert eng.sm2dsa_verify(q, e, rs) == exp
eng.close()
print("sanitize pass ok")
