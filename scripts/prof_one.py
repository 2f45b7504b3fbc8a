"""One launch of one batch kernel on device-resident synthetic inputs (for ncu).
usage: prof_one.py <curve> <verify|mul_var|mul_gen> <log2 n> [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ecb200

curve, op, lg = sys.argv[1], sys.argv[2], int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
n = 1 << lg
fb = ecb200.field_bytes(curve)
eng = ecb200.Engine(0)
dev = torch.device("cuda:0")
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
st = ts.cuda_stream
rng = np.random.default_rng(7)


def rnd(cols):
    a = rng.integers(0, 256, size=(n, cols), dtype=np.uint8)
    a[:, 0] &= 0x3F
    return torch.from_numpy(a).to(dev)


slot = 1 + 2 * fb
ks = rnd(fb)
pts_slots = torch.empty(n * slot, dtype=torch.uint8, device=dev)
eng.mul_gen_dev(curve, n, ks, pts_slots, ecb200.FLAG_UNCOMPRESSED, st)
pts = pts_slots.view(n, slot)[:, 1:].contiguous()
k2 = rnd(fb)
out = torch.empty(n * slot, dtype=torch.uint8, device=dev)
z = rnd(fb)
rs = rnd(2 * fb)
rs.view(n, 2 * fb)[:, fb] &= 0x3F
ok = torch.empty(n, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps):
    e0.record()
    if op == "verify":
        eng.ecdsa_verify_dev(curve, n, pts, z, rs, ok, st)
    elif op == "mul_var":
        eng.mul_var_dev(curve, n, pts, None, k2, out, None, ecb200.FLAG_UNCOMPRESSED, st)
    elif op == "mul_var_ct":
        eng.mul_var_dev(curve, n, pts, None, k2, out, None, ecb200.FLAG_UNCOMPRESSED | ecb200.FLAG_CT, st)
    else:
        eng.mul_gen_dev(curve, n, ks, out, ecb200.FLAG_UNCOMPRESSED | ecb200.FLAG_CT, st)
    e1.record()
    torch.cuda.synchronize()
    print(f"{curve} {op} n={n}: {e0.elapsed_time(e1):.3f} ms  {n / e0.elapsed_time(e1) / 1e3:.3f} M/s")
