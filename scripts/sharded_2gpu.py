"""torchrun check of the multi-GPU host logic on real GPUs (NCCL): verify_sharded + lincomb_sharded against the oracle.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/sharded_2gpu.py"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ecb200
from oracle import ecoracle as o
import importlib
sh = importlib.import_module("rustcrypto-elliptic-curves_b200.sharding")

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
eng = ecb200.Engine(local)
for cname in ("k256", "p256"):
    c = o.curve(cname)
    rng = random.Random(5)            # same inputs on every rank
    n = 37
    pts = [o.mul_gen(c, rng.randrange(1, c.n)) for _ in range(n)]
    ks = [rng.randrange(c.n) for _ in range(n)]
    pb = b"".join(P[0].to_bytes(c.fb, "big") + P[1].to_bytes(c.fb, "big") for P in pts)
    kb = b"".join(k.to_bytes(c.fb, "big") for k in ks)
    got = sh.lincomb_sharded(eng, cname, pb, kb)
    exp = o.slot_encode(c, o.pt_lincomb(c, list(zip(pts, ks))))
    assert got == exp, (rank, cname)
    # verify_sharded with gather
    from tests import nextrows
    q, z, rs = bytearray(), bytearray(), bytearray()
    for i in range(41):
        d, k, zz, (r, s, _) = nextrows.make_sig(c, rng)
        Q = o.mul_gen(c, d)
        if i % 4 == 1:
            s ^= 1
        q += Q[0].to_bytes(c.fb, "big") + Q[1].to_bytes(c.fb, "big"); z += zz; rs += r.to_bytes(c.fb, "big") + s.to_bytes(c.fb, "big")
    mask = sh.verify_sharded(eng, cname, bytes(q), bytes(z), bytes(rs), gather=True)
    if rank == 0:
        assert mask == o.batch_verify(c, bytes(q), bytes(z), bytes(rs)), cname
dist.barrier()
if rank == 0:
    print("sharded %d-GPU NCCL check ok" % world)
dist.destroy_process_group()
