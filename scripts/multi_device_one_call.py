"""ONE process, ONE context over all the GPUs of the box (ecb200_init_multi): the library shards every host-pointer call by
contiguous index range, one host thread + streams + pinned staging per device.  Times BASELINE configs[2] / [3] (2^22 ECDSA
verifications in ONE call) end to end - host buffers in, mask in a host buffer out - for 1, 2, 4, ... devices and prints one
JSON line per device count.  usage: multi_device_one_call.py [curve=k256] [log2 rows=22] [steps=5]"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ecb200
from ecb200 import workloads as wl

curve = sys.argv[1] if len(sys.argv) > 1 else "k256"
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 22
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n = 1 << lg
ndev = torch.cuda.device_count()
e0 = ecb200.Engine(0)
q, z, rs, exp = wl.make_verify_batch(wl.EngineBackend(e0, curve), curve, n, 0xB2000003)
e0.close()
pin = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (q, z, rs)]
out = torch.empty(n, dtype=torch.uint8).pin_memory()
ptr = [ctypes.c_void_p(t.data_ptr()) for t in pin]
o = ctypes.c_void_p(out.data_ptr())
cid = ecb200.curve_id(curve)
base = None
g = 1
while g <= ndev:
    m = ecb200.Engine(devices=list(range(g)))
    for _ in range(2):
        assert m.lib.ecb200_ecdsa_verify(m.h, cid, n, ptr[0], ptr[1], ptr[2], o) == 0
    assert np.array_equal(out.numpy(), exp), "mask differs from the construction at %d devices" % g
    t0 = time.perf_counter()
    for _ in range(steps):
        m.lib.ecb200_ecdsa_verify(m.h, cid, n, ptr[0], ptr[1], ptr[2], o)
    dt = (time.perf_counter() - t0) / steps
    v = n / dt
    base = base or v
    print(json.dumps({"what": "one process, one ecb200_init_multi context, one ecb200_ecdsa_verify call per step (host pointers, page-locked)",
                      "curve": curve, "rows": n, "devices": g, "ms_per_call": round(dt * 1e3, 3), "verifies_per_s": round(v, 1),
                      "speedup_vs_1_device": round(v / base, 3), "mask_ok": True}), flush=True)
    m.close()
    g *= 2
