"""One launch of a secret-scalar kernel with a chosen SECRET pattern and fixed public inputs (for scripts/ct_audit.sh: the
ncu counters of the launches must not depend on the secrets).  usage: ct_dyn_run.py <curve> <mul_var|mul_gen|sign>
The operation runs once per secret pattern - small (1, 2, 3, ...), random, high (n-1, n-2, ...), sparse (single bits) - each in
its own cudaProfilerStart/Stop range, in this order, then `small` once more (identical inputs: any counter that differs between
the two `small` runs is measurement noise, not data dependence); scripts/ct_audit_compare.py splits the ncu launch list into five equal parts."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ecb200
from oracle import ecoracle as o

curve, op = sys.argv[1], sys.argv[2]
PATTERNS = ["small", "random", "high", "sparse", "small_again"]      # the last one repeats the first: measurement-noise control
c = o.curve(curve)
fb = c.fb
n = 1 << 14
eng = ecb200.Engine(0)
dev = torch.device("cuda:0")
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
st = ts.cuda_stream


def secrets(pattern, salt):
    pattern = "small" if pattern == "small_again" else pattern
    if pattern == "random":
        a = np.random.default_rng(100 + salt).integers(0, 256, size=(n, fb), dtype=np.uint8)
        a[:, 0] &= 0x7F
        a[:, -1] |= 1
        return a
    vals = {"small": lambda i: i + 1 + salt, "high": lambda i: c.n - 1 - i - salt, "sparse": lambda i: 1 << ((i * 7 + salt) % (8 * fb - 2))}[pattern]
    return np.frombuffer(b"".join(int(vals(i)).to_bytes(fb, "big") for i in range(n)), np.uint8).reshape(n, fb).copy()


pub = np.random.default_rng(7).integers(0, 256, size=(n, fb), dtype=np.uint8)
pub[:, 0] &= 0x7F
slot = 1 + 2 * fb
pts_slots = torch.empty(n * slot, dtype=torch.uint8, device=dev)
eng.mul_gen_dev(curve, n, torch.from_numpy(pub).to(dev), pts_slots, ecb200.FLAG_UNCOMPRESSED, st)
pts = pts_slots.view(n, slot)[:, 1:].contiguous()
z = torch.from_numpy(np.random.default_rng(8).integers(0, 256, size=(n, fb), dtype=np.uint8)).to(dev)
out = torch.empty(n * slot, dtype=torch.uint8, device=dev)
rs = torch.empty(n * 2 * fb, dtype=torch.uint8, device=dev)
rid = torch.empty(n, dtype=torch.uint8, device=dev)
ok = torch.empty(n, dtype=torch.uint8, device=dev)
for pattern in PATTERNS:
    k = torch.from_numpy(secrets(pattern, 0)).to(dev)
    d = torch.from_numpy(secrets(pattern, 3)).to(dev)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    if op == "mul_var":
        eng.mul_var_dev(curve, n, pts, None, k, out, None, ecb200.FLAG_UNCOMPRESSED | ecb200.FLAG_CT, st)
    elif op == "mul_gen":
        eng.mul_gen_dev(curve, n, k, out, ecb200.FLAG_UNCOMPRESSED | ecb200.FLAG_CT, st)
    else:
        eng.ecdsa_sign_dev(curve, n, d, k, z, rs, rid, ok, st)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(curve, op, pattern, "ok rows:", int(ok.sum().item()) if op == "sign" else "-")
