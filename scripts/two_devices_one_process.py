"""One process, one context per GPU (the C ABI allows it; the benchmark uses one process per GPU): every curve's fixed-base
kernel (which needs the > 48 KB shared-memory opt-in on P-384) and the verify path must work on each device."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ecb200
from oracle import ecoracle as o
from tests import nextrows

n_dev = torch.cuda.device_count()
engs = [ecb200.Engine(d) for d in range(n_dev)]
rng = random.Random(3)
for cname in ("p384", "k256", "p256", "sm2"):
    c = o.curve(cname)
    ks = b"".join(rng.randrange(c.n).to_bytes(c.fb, "big") for _ in range(50))
    exp = o.batch_mul_gen(c, ks)
    q, z, rs = bytearray(), bytearray(), bytearray()
    for i in range(20):
        d, k, zz, (r, s, _) = nextrows.make_sig(c, rng)
        Q = o.mul_gen(c, d)
        q += Q[0].to_bytes(c.fb, "big") + Q[1].to_bytes(c.fb, "big"); z += zz; rs += r.to_bytes(c.fb, "big") + (s ^ (i & 1)).to_bytes(c.fb, "big")
    expv = o.batch_verify(c, bytes(q), bytes(z), bytes(rs))
    for e in engs:
        assert e.mul_by_generator_batch(cname, ks, ecb200.FLAG_CT) == exp, (cname, e.device)
        assert e.ecdsa_verify(cname, bytes(q), bytes(z), bytes(rs)) == expv, (cname, e.device)
for e in engs:
    e.close()
print("%d devices in one process ok" % n_dev)
