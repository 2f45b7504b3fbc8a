"""Multi-GPU host logic: one process per GPU, batches shard by contiguous index range, no
collective on the data path.  The only exchange is the optional multi-point lincomb, whose <= 8 partial sums
(3*FB bytes each) are all-gathered and added by every rank.

`engine` is an ecb200.Engine (or anything with the same ecdsa_verify / mul_batch / lincomb methods — the CPU
gloo tests plug in an oracle-backed stand-in because there is no GPU in the build container)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import FLAG_PROJ, field_bytes, shard_range


def _rank_world(group=None) -> Tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _xdev(group=None):
    """Device the collective's tensors must live on: the current CUDA device under NCCL, the host under gloo."""
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def verify_sharded(engine, curve, q: bytes, z: bytes, rs: bytes, group=None, gather: bool = False) -> Optional[bytes]:
    """Each rank verifies rows [lo, hi) of the GLOBAL batch (every rank holds, or can address, the whole batch).
    Returns this rank's ok bytes; with gather=True rank 0 returns the concatenated mask of all ranks (a control-plane
    gather of 1 byte per row, outside the timed data path) and the other ranks return None."""
    import torch
    import torch.distributed as dist
    fb = field_bytes(curve)
    n = len(z) // fb
    rank, world = _rank_world(group)
    lo, hi = shard_range(n, rank, world)
    ok = engine.ecdsa_verify(curve, q[2 * fb * lo:2 * fb * hi], z[fb * lo:fb * hi], rs[2 * fb * lo:2 * fb * hi])
    if not gather or world == 1:
        return ok
    sizes = [shard_range(n, r, world) for r in range(world)]
    mx = max(h - l for l, h in sizes)
    dev = _xdev(group)
    mine = torch.zeros(mx, dtype=torch.uint8)
    mine[:hi - lo] = torch.frombuffer(bytearray(ok), dtype=torch.uint8)
    mine = mine.to(dev)
    bufs = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return b"".join(bytes(bufs[r][:h - l].cpu().numpy().tobytes()) for r, (l, h) in enumerate(sizes))


def lincomb_sharded(engine, curve, pts: bytes, ks: bytes, group=None, flags: int = 0) -> bytes:
    """sum_i k_i P_i over terms sharded by index range: each rank reduces its terms to one projective partial on its
    GPU, the partials (3*FB bytes per rank) are all-gathered, and every rank adds them (scalars = 1, projective
    inputs) to the same SEC1 slot.  LinearCombinationExt over a slice: k256/src/arithmetic/mul.rs:326-340."""
    import torch
    import torch.distributed as dist
    fb = field_bytes(curve)
    n = len(ks) // fb
    rank, world = _rank_world(group)
    lo, hi = shard_range(n, rank, world)
    partial = engine.lincomb(curve, pts[2 * fb * lo:2 * fb * hi], ks[fb * lo:fb * hi], flags, True)   # X||Y||Z
    if world == 1:
        parts = partial
    else:
        dev = _xdev(group)
        mine = torch.frombuffer(bytearray(partial), dtype=torch.uint8).to(dev)
        bufs = [torch.zeros(3 * fb, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(bufs, mine, group=group)      # NCCL over NVLink on the GPU box, gloo in the CPU tests
        parts = b"".join(bytes(b.cpu().numpy().tobytes()) for b in bufs)
    ones = ((1).to_bytes(fb, "big")) * world
    return engine.lincomb(curve, parts, ones, flags | FLAG_PROJ, False)
