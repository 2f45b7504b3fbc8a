"""One invocation of one batch operation on device-resident synthetic inputs, bracketed by cudaProfilerStart/Stop (for
`ncu --profile-from-start off`: every kernel of the operation, nothing of the set-up).
usage: prof_one.py <curve> <verify|verify_keys|mul_var|mul_var_ct|mul_var_proj|mul_var_proj_ct|mul_gen|sign> <log2 n> [reps]
verify = every row its own key (per-row path); verify_keys = 2^16 distinct keys reused round-robin (BASELINE configs 2 / 3: per-key tables)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ecb200

curve, op, lg = sys.argv[1], sys.argv[2], int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
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
if op == "verify_keys":
    pts = pts[torch.arange(n, device=dev) % min(n, 1 << 16)].contiguous()
    op = "verify"
k2 = rnd(fb)
out = torch.empty(n * slot, dtype=torch.uint8, device=dev)
z = rnd(fb)
rs = rnd(2 * fb)
rs.view(n, 2 * fb)[:, fb] &= 0x3F
ok = torch.empty(n, dtype=torch.uint8, device=dev)
rid = torch.empty(n, dtype=torch.uint8, device=dev)
xyz = None
if "proj" in op:      # X = x l, Y = y l, Z = l with a random l (BASELINE config 2 shape), made with the engine's field hook
    torch.cuda.synchronize()
    lam = rnd(fb).cpu().numpy()
    p = pts.cpu().numpy()
    X, _ = eng.field_op(curve, 0, 2, np.ascontiguousarray(p[:, :fb]), lam)
    Y, _ = eng.field_op(curve, 0, 2, np.ascontiguousarray(p[:, fb:]), lam)
    xyz = torch.from_numpy(np.concatenate([np.frombuffer(X, np.uint8).reshape(n, fb), np.frombuffer(Y, np.uint8).reshape(n, fb), lam], axis=1)).to(dev)
CT, U, PR = ecb200.FLAG_CT, ecb200.FLAG_UNCOMPRESSED, ecb200.FLAG_PROJ


def run():
    if op == "verify":
        eng.ecdsa_verify_dev(curve, n, pts, z, rs, ok, st)
    elif op == "mul_var":
        eng.mul_var_dev(curve, n, pts, None, k2, out, None, U, st)
    elif op == "mul_var_ct":
        eng.mul_var_dev(curve, n, pts, None, k2, out, None, U | CT, st)
    elif op == "mul_var_proj":
        eng.mul_var_dev(curve, n, xyz, None, k2, out, None, U | PR, st)
    elif op == "mul_var_proj_ct":
        eng.mul_var_dev(curve, n, xyz, None, k2, out, None, U | PR | CT, st)
    elif op == "sign":
        eng.ecdsa_sign_dev(curve, n, ks, k2, z, rs, rid, ok, st)
    else:
        eng.mul_gen_dev(curve, n, ks, out, U | CT, st)


run()      # warm-up: builds the big fixed-base table on the first verify, sets function attributes
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
for r in range(reps):
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{curve} {op} n={n}: {e0.elapsed_time(e1):.3f} ms  {n / e0.elapsed_time(e1) / 1e3:.3f} M/s")
torch.cuda.cudart().cudaProfilerStop()
