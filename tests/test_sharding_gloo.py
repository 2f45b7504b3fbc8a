"""world_size-2 gloo test of the multi-GPU host logic (index sharding, partial-sum exchange for lincomb), on CPU.
The engine is replaced by an oracle-backed stand-in with the same method signatures: this checks the plumbing
(ranges, gather order, projective partial format), not the kernels."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ecoracle as o


class OracleEngine:
    """Same call surface as ecb200.Engine for the methods sharding.py uses."""

    def ecdsa_verify(self, curve, q, z, rs):
        return o.batch_verify(o.curve(curve), q, z, rs)

    def lincomb(self, curve, points, ks, flags=0, out_proj=False):
        c = o.curve(curve)
        fb = c.fb
        n = len(ks) // fb
        terms = []
        for i in range(n):
            k = int.from_bytes(ks[i * fb:(i + 1) * fb], "big")
            if flags & 8:
                X, Y, Z = (int.from_bytes(points[(3 * i + j) * fb:(3 * i + j + 1) * fb], "big") for j in range(3))
                P = o.proj_to_affine(c, X, Y, Z)
            else:
                P = (int.from_bytes(points[2 * i * fb:(2 * i + 1) * fb], "big"), int.from_bytes(points[(2 * i + 1) * fb:(2 * i + 2) * fb], "big"))
            terms.append((P, k))
        R = o.pt_lincomb(c, terms)
        if out_proj:
            if R is None:
                return (0).to_bytes(fb, "big") + (1).to_bytes(fb, "big") + (0).to_bytes(fb, "big")
            lam = 7
            return (R[0] * lam % c.p).to_bytes(fb, "big") + (R[1] * lam % c.p).to_bytes(fb, "big") + lam.to_bytes(fb, "big")
        return o.slot_encode(c, R)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cname, q, z, rs, pts, ks, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    sh = importlib.import_module("rustcrypto-elliptic-curves_b200.sharding")
    eng = OracleEngine()
    mask = sh.verify_sharded(eng, cname, q, z, rs, gather=True)
    point = sh.lincomb_sharded(eng, cname, pts, ks)
    if rank == 0:
        ret["mask"] = mask
    ret["point%d" % rank] = point
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cname", ["k256", "p256"])
def test_two_rank_sharding(cname):
    c = o.curve(cname)
    fb = c.fb
    rng = random.Random(3)
    rows = []
    for i in range(7):     # odd count: uneven shards
        d = rng.randrange(1, c.n)
        Q = o.mul_gen(c, d)
        zb = rng.randrange(1 << 8 * fb).to_bytes(fb, "big")
        k = rng.randrange(1, c.n)
        r = o.mul_gen(c, k)[0] % c.n
        s = pow(k, -1, c.n) * (o.reduce_once(c, int.from_bytes(zb, "big")) + r * d) % c.n
        if c.low_s and s > c.n >> 1:
            s = c.n - s
        if i % 3 == 1:
            s ^= 8
        rows.append((Q, zb, r, s))
    q = b"".join(P[0].to_bytes(fb, "big") + P[1].to_bytes(fb, "big") for P, _, _, _ in rows)
    z = b"".join(r[1] for r in rows)
    rs = b"".join(r[2].to_bytes(fb, "big") + r[3].to_bytes(fb, "big") for r in rows)
    pts_l = [o.mul_gen(c, rng.randrange(1, 1 << 40)) for _ in range(5)]
    ks_l = [rng.randrange(c.n) for _ in range(5)]
    pts = b"".join(P[0].to_bytes(fb, "big") + P[1].to_bytes(fb, "big") for P in pts_l)
    ks = b"".join(k.to_bytes(fb, "big") for k in ks_l)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, cname, q, z, rs, pts, ks, ret), nprocs=2, join=True)
    assert ret["mask"] == o.batch_verify(c, q, z, rs)
    exp = o.slot_encode(c, o.pt_lincomb(c, list(zip(pts_l, ks_l))))
    assert ret["point0"] == exp and ret["point1"] == exp
