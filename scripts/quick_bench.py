"""Quick kernel-only timing of every batch kernel (device-resident inputs, CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ecb200
from oracle import ecoracle as o

eng = ecb200.Engine(0)
dev = torch.device("cuda:0")
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
stream = ts.cuda_stream
assert stream != 0


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def rand_scalars(n, fb, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, fb), dtype=np.uint8)
    a[:, 0] &= 0x7F
    return a


# usage: quick_bench.py [curve,curve,...] [log2 n]   (default: every curve at its quick size)
CASES = (("k256", 1 << 18), ("p256", 1 << 17), ("sm2", 1 << 17), ("p384", 1 << 16), ("p192", 1 << 17), ("p224", 1 << 17))
if len(sys.argv) > 1:
    CASES = tuple((c, (1 << int(sys.argv[2])) if len(sys.argv) > 2 else dict(CASES)[c]) for c in sys.argv[1].split(","))
for cname, n in CASES:
    c = o.curve(cname)
    fb = c.fb
    ks = torch.from_numpy(rand_scalars(n, fb, 1)).to(dev)
    # points: k*G computed by the engine itself (uncompressed), then strip the tag
    slot = 1 + 2 * fb
    out = torch.empty(n * slot, dtype=torch.uint8, device=dev)
    t = timeit(lambda: eng.mul_gen_dev(cname, n, ks, out, ecb200.FLAG_UNCOMPRESSED | ecb200.FLAG_CT, stream))
    print(f"{cname} mul_gen CT      n={n}: {t:8.3f} ms  {n / t / 1e3:8.3f} M/s", flush=True)
    t = timeit(lambda: eng.mul_gen_dev(cname, n, ks, out, ecb200.FLAG_UNCOMPRESSED, stream))
    print(f"{cname} mul_gen vartime n={n}: {t:8.3f} ms  {n / t / 1e3:8.3f} M/s", flush=True)
    pts = out.view(n, slot)[:, 1:].contiguous()
    k2 = torch.from_numpy(rand_scalars(n, fb, 2)).to(dev)
    out2 = torch.empty(n * slot, dtype=torch.uint8, device=dev)
    for ct in (1, 0):
        t = timeit(lambda: eng.mul_var_dev(cname, n, pts, None, k2, out2, None, ecb200.FLAG_UNCOMPRESSED | ct, stream))
        print(f"{cname} mul_var ct={ct}    n={n}: {t:8.3f} ms  {n / t / 1e3:8.3f} M/s", flush=True)
    # verify: random (invalid) signatures exercise the full arithmetic path
    z = torch.from_numpy(rand_scalars(n, fb, 3)).to(dev)
    rs = torch.from_numpy(rand_scalars(n, 2 * fb, 4)).to(dev)
    rs.view(n, 2 * fb)[:, fb] &= 0x3F
    ok = torch.empty(n, dtype=torch.uint8, device=dev)
    t = timeit(lambda: eng.ecdsa_verify_dev(cname, n, pts, z, rs, ok, stream))
    print(f"{cname} verify          n={n}: {t:8.3f} ms  {n / t / 1e3:8.3f} M/s  (accepted {int(ok.sum())})", flush=True)
print("launches", eng.launch_count)
