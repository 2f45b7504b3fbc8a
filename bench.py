#!/usr/bin/env python3
"""bench.py — headline benchmark of the batched EC hot path on B200 (contract: see the task prompt).

    python bench.py --gpus N --steps K --warmup W          # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N ...           # the reference's CPU algorithms (C++ port) on host cores

Headline workload (BASELINE.json configs[2], the config the metric "ECDSA verify/s at 1/2/4/8 B200" is quoted
on; it fits one GPU): secp256k1 ECDSA verify_prehash over 2^22 synthetic signatures per GPU, 2^16 distinct
keys, 1/16 of the rows corrupted, sharded by index range with no collective on the data path (weak scaling:
every rank verifies its own 2^22 rows).  A "step" is one pass over the rank's batch.
  value  = whole-job verifies/s with inputs resident in HBM (CUDA events on the launch stream, max over ranks)
  e2e    = the same through the host-pointer C-ABI call (H2D of 160 B/row and D2H of 1 B/row inside the timing)
  roofline: INT32 multiply-add issue rate — achieved = rows/s x W_elem (reference-algorithm
            32x32->64 products per row) / the dominant kernel's launch duration (CUDA events inside the library) against
            the measured whole-chip IMAD.WIDE rate (peaks_int.json);
            an HBM figure (algorithmic bytes / time vs MEASURED_PEAKS.json) is reported beside it; executed_frac /
            fmaheavy_pipe_pct (ncu, profiles/summary.json) are the hardware-utilisation view.
  cpu_baseline: oracle/ecport.cpp (C++ port of the reference algorithms, OpenMP) on a bounded sample.
The other BASELINE configs are measured in the same run at N = 1 and reported under "others".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_PER_M = {"k256": 73, "p256": 136, "sm2": 136, "p384": 300}
M_REF = {"verify_k256": 3218, "verify_p256": 9005, "mul_gen_k256": 817, "mul_var_k256": 1990,
         "mul_var_p384": 6478, "mul_var_sm2": 4366, "mul_var_p256": 4366}
IO_BYTES = {"verify": lambda fb: 5 * fb + 1, "mul_gen": lambda fb: fb + 1 + fb, "mul_var": lambda fb: 4 * fb + 1 + fb}


def load_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) >= 7 and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def imad_peak_gmacs(sm_mhz):
    """Measured IMAD.WIDE issue peak: whole-chip multiply-accumulates per second of IMAD.WIDE.U32.X carry chains
    (bench/imad_peak.cu, event-timed on this pool's B200 at 1965 MHz -> peaks_int.json), scaled by the SM clock seen
    during the timed region when that is lower."""
    pk = load_json(os.path.join(ROOT, "peaks_int.json"), {})
    g = pk.get("imad_wide_chip_gmacs")
    src = "measured (peaks_int.json: IMAD.WIDE.U32.X carry chains, whole chip, CUDA events)"
    if not g:
        g, src = 32.0 * 148 * 1965e6 / 1e9, "fallback: 32 IMAD.WIDE/clk/SM x 148 SMs x 1965 MHz"
    ref_mhz = pk.get("sm_mhz_during_measurement", 1965.0)
    mhz = sm_mhz or ref_mhz
    return g * min(1.0, mhz / ref_mhz), src


def roofline(kind, curve, rows_per_s, ms_kernel, n_rows, sm_mhz, kernel_ms=None, kernel_name=None):
    """achieved = algorithmic multiply-accumulates of one launch (reference-algorithm field
    multiplications x schoolbook 32x32 products each) / launch duration; kernel_ms (CUDA events around the dominant
    kernel, ecb200_kernel_timing) is used when available, else the whole step."""
    fb = 48 if curve == "p384" else 32
    w_elem = M_REF[f"{kind}_{curve}"] * W_PER_M[curve]
    peak, src = imad_peak_gmacs(sm_mhz)
    dur = kernel_ms if kernel_ms else ms_kernel
    ach = n_rows * w_elem / (dur * 1e-3) / 1e9
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {})
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = n_rows * IO_BYTES[kind](fb) / (ms_kernel * 1e-3) / 1e9
    prof = load_json(os.path.join(ROOT, "profiles", "summary.json"), {}).get(f"{kind}_{curve}", {})
    out = {"bound": "imad", "achieved": round(ach, 1), "peak": round(peak, 1), "unit": "Gmac/s (32x32->64 multiply-accumulates)",
           "frac": round(ach / peak, 4), "w_elem": w_elem, "peak_source": src,
           "kernel": kernel_name, "kernel_ms_per_launch": round(dur, 4), "rows_per_launch": n_rows,
           "note": "achieved counts the REFERENCE algorithm's multiplications (M_ref x W_per_M); the device executes fewer, "
                   "so this throughput-normalised fraction can exceed 1 - executed_frac and fmaheavy_pipe_pct are the hardware view",
           "traffic": prof.get("dram_bytes_per_launch"), "traffic_rows": prof.get("n_rows"),
           "hbm": {"achieved": round(hbm_ach, 2), "peak": hbm_peak, "unit": "GB/s", "frac": round(hbm_ach / hbm_peak, 5),
                   "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"}}
    if prof.get("wide_macs_per_row"):
        ex = n_rows * prof["wide_macs_per_row"] / (dur * 1e-3) / 1e9
        out["executed"] = round(ex, 1)
        out["executed_frac"] = round(ex / peak, 4)
        out["fmaheavy_pipe_pct"] = prof.get("pipe_fmaheavy_pct")
        out["executed_source"] = prof.get("source")
    return out


# ------------------------------------------------------------------------------------------------ reference arm
class PyBackend:
    """Input generation for the reference arm without the CUDA engine: k*G from the C++ port, mod-n algebra in
    Python integers (bounded sample sizes only)."""

    def __init__(self, curve):
        from oracle import ecoracle as o
        from tests import port_lib
        self.c, self.lib = o.curve(curve), port_lib.load()
        self.cid, self.fb = self.c.cid, self.c.fb

    def mul_gen_xy(self, k):
        n = k.shape[0]
        out = (ctypes.c_uint8 * (n * (1 + 2 * self.fb)))()
        self.lib.port_mul_gen(self.cid, n, k.tobytes(), out, 0)
        return np.frombuffer(bytes(out), np.uint8).reshape(n, 1 + 2 * self.fb)[:, 1:]

    def fn(self, op, a, b=None):
        n_, fb = self.c.n, self.fb
        A = [int.from_bytes(r.tobytes(), "big") for r in a]
        B = [int.from_bytes(r.tobytes(), "big") for r in b] if b is not None else A
        f = {0: lambda x, y: (x + y) % n_, 2: lambda x, y: x * y % n_, 4: lambda x, y: (-x) % n_, 5: lambda x, y: pow(x, -1, n_) if x % n_ else 0}[op]
        return np.frombuffer(b"".join(f(x, y).to_bytes(fb, "big") for x, y in zip(A, B)), np.uint8).reshape(a.shape).copy()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import port_lib
    wl = importlib_pkg().workloads
    lib = port_lib.load()
    lib.port_set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    cores = lib.port_threads()
    curve = "k256"
    # bounded sample: sized for roughly 2 s per step on this host
    n0 = 4096
    q, z, rs, exp = wl.make_verify_batch(PyBackend(curve), curve, n0, 0xB2000003)
    ok = (ctypes.c_uint8 * n0)()
    t = time.time()
    lib.port_verify(0, n0, q.tobytes(), z.tobytes(), rs.tobytes(), ok)
    rate = n0 / max(time.time() - t, 1e-6)
    assert bytes(ok) == exp.tobytes(), "C++ port disagrees with the constructed mask"
    n = int(min(1 << 20, max(n0, rate * 2.0)))
    reps = (n + n0 - 1) // n0
    qq, zz, rr = np.tile(q, (reps, 1))[:n], np.tile(z, (reps, 1))[:n], np.tile(rs, (reps, 1))[:n]
    qb, zb, rb = qq.tobytes(), zz.tobytes(), rr.tobytes()
    okn = (ctypes.c_uint8 * n)()
    for _ in range(args.warmup):
        lib.port_verify(0, n, qb, zb, rb, okn)
    t = time.time()
    for _ in range(args.steps):
        lib.port_verify(0, n, qb, zb, rb, okn)
    dt = time.time() - t
    value = n * args.steps / dt
    sample = f"{n} rows per step ({n0} distinct signatures tiled), {args.steps} steps"
    emit({
        "impl": "reference", "metric": "secp256k1 ECDSA verify_prehash throughput", "value": round(value, 1), "unit": "verifies/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (5x52 field, u128 accumulators)",
        "data": "synthetic", "config": config_block(args.gpus, bounded=sample),
        "cpu_baseline": {"value": round(value, 1), "unit": "verifies/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 1), "unit": "verifies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = C++ port of the reference's CPU algorithms (oracle/ecport.cpp, OpenMP); the Rust reference cannot be built here (no cargo/rustc)"})


def importlib_pkg():
    import importlib
    pkg = importlib.import_module("rustcrypto-elliptic-curves_b200")
    importlib.import_module("rustcrypto-elliptic-curves_b200.workloads")
    return pkg


def config_block(n_gpus, bounded=None, curve="k256", log2_rows=22):
    rows = 1 << log2_rows
    if curve == "k256":
        what = "BASELINE configs[2]: secp256k1 ECDSA verify_prehash, 2^%d signatures per GPU" % log2_rows
    else:   # --curve p256: BASELINE configs[3] (P-256 verify via the primeorder path), same construction, same sharding
        what = "BASELINE configs[3]: p256 ECDSA verify_prehash, 2^%d signatures per GPU" % log2_rows
    c = {"workload": what + " (2^16 distinct keys, 1/16 rows corrupted), sharded by index range, no collective",
         "rows_per_gpu": rows, "global_rows": rows * n_gpus, "curve": curve,
         "l2_policy": "inputs (%d MB per step) larger than L2 (126 MB)" % (rows * 160 // 1000000), "parallelism": f"index-shard x{n_gpus}"}
    if bounded:
        c["bounded_sample"] = bounded
    return c


# ------------------------------------------------------------------------------------------------ our arm
_REAL_STDOUT = None


def quiet_stdout():
    """Rank 0 must print ONE JSON line: route everything libraries write to fd 1 (NCCL's version banner, torchrun
    chatter) to stderr and keep a private handle on the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--log2-rows", type=int, default=22, help="rows per GPU (default 2^22 = the BASELINE config)")
    ap.add_argument("--curve", default="k256", choices=["k256", "p256"], help="k256 = BASELINE configs[2] (the headline), p256 = configs[3]")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib_pkg()
    wl = pkg.workloads
    eng = pkg.Engine(local)
    dev = torch.device("cuda", local)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    curve, n = args.curve, 1 << args.log2_rows
    fb = 32
    seed = 0xB2000003 if curve == "k256" else 0xB2000004

    # ---- inputs (synthetic, generated by the engine; not timed)
    q, z, rs, exp = wl.make_verify_batch(wl.EngineBackend(eng, curve), curve, n, seed + 1000 * rank)
    d_q, d_z, d_rs = torch.from_numpy(q).to(dev), torch.from_numpy(z).to(dev), torch.from_numpy(rs).to(dev)
    d_ok = torch.empty(n, dtype=torch.uint8, device=dev)
    h_ok = np.empty(n, np.uint8)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- kernel-resident timing
    for _ in range(max(args.warmup, 3)):
        eng.ecdsa_verify_dev(curve, n, d_q, d_z, d_rs, d_ok, st)
    barrier()
    assert np.array_equal(d_ok.cpu().numpy(), exp), "verify mask differs from the constructed expectation"
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.kernel_timing(True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record(ts)
        for _ in range(args.steps):
            eng.ecdsa_verify_dev(curve, n, d_q, d_z, d_rs, d_ok, st)
        e1.record(ts)
        barrier()
    eng.kernel_timing(False)
    k_ms, k_cnt = eng.kernel_timing_read()
    kernel_ms = max_over_ranks(k_ms / max(k_cnt, 1))
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count - l0
    ms_step = ms_total / args.steps
    value = n * world / (ms_step * 1e-3)
    clocks = clk.summary()

    # ---- end to end through the host-pointer C ABI (H2D + kernels + D2H per step)
    # inputs and the result live in page-locked host memory (the contract's "from pinned host memory"); the library
    # copies chunk i+1 H2D and chunk i-1 D2H on side streams while chunk i computes
    pq, pz, prs = (torch.from_numpy(a).pin_memory() for a in (q, z, rs))
    pok = torch.empty(n, dtype=torch.uint8).pin_memory()
    hq, hz, hrs, hok = pq.numpy(), pz.numpy(), prs.numpy(), pok.numpy()
    for _ in range(2):
        eng.ecdsa_verify(curve, hq, hz, hrs, out=hok)
    assert np.array_equal(hok, exp)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.ecdsa_verify(curve, hq, hz, hrs, out=hok)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = n * world * args.steps / dt

    out = {
        "metric": ("secp256k1" if curve == "k256" else "P-256") + " ECDSA verify_prehash throughput", "value": round(value, 1), "unit": "verifies/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (8x32, IMAD.WIDE carry chains)", "data": "synthetic",
        "config": config_block(world, None, curve, args.log2_rows), "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": round(e2e_value, 1), "unit": "verifies/s", "h2d_bytes_per_step": int(n * 5 * fb), "d2h_bytes_per_step": int(n),
                "ms_per_step": round(dt / args.steps * 1e3, 3), "api": "ecb200_ecdsa_verify (host pointers to page-locked buffers; chunked H2D / compute / D2H overlap inside the call)"},
        "roofline": roofline("verify", curve, value / world, ms_step, n, clocks.get("sm_mhz"), kernel_ms,
                             "k_verify_main<CurveK256, VM_ECDSA>" if curve == "k256" else "k_verify_main<CurveP256, VM_ECDSA> (+ k_wintab<CurveP256> before it)"),
    }

    if rank == 0 and world == 1 and not args.no_others and curve == "k256":
        out["others"] = other_configs(pkg, eng, dev, ts)
    if rank == 0 and world == 1 and not args.no_cpu and curve == "k256":
        out["cpu_baseline"] = cpu_baseline(q, z, rs, exp)
    if rank == 0:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(q, z, rs, exp):
    """C++ port of the reference algorithms on the host cores, bounded sample of the same batch."""
    from tests import port_lib
    lib = port_lib.load()
    n0 = 1 << 12
    ok = (ctypes.c_uint8 * n0)()
    t = time.time()
    lib.port_verify(0, n0, q[:n0].tobytes(), z[:n0].tobytes(), rs[:n0].tobytes(), ok)
    rate = n0 / max(time.time() - t, 1e-6)
    n = int(min(q.shape[0], max(n0, rate * 8.0)))
    ok = (ctypes.c_uint8 * n)()
    qb, zb, rb = q[:n].tobytes(), z[:n].tobytes(), rs[:n].tobytes()
    t = time.time()
    lib.port_verify(0, n, qb, zb, rb, ok)
    dt = time.time() - t
    agree = bytes(ok) == exp[:n].tobytes()
    out = {"value": round(n / dt, 1), "unit": "verifies/s", "cores": lib.port_threads(), "kind": "port",
           "sample": f"first {n} rows of the same batch, one pass ({dt:.1f} s)", "matches_gpu_mask": bool(agree)}
    try:   # second CPU figure: OpenSSL libcrypto (ECDSA_do_verify + low-s rule) on the same cores
        from oracle import libcrypto_ref as lc
        m = 1 << 13
        t = time.time()
        lc.verify_batch("k256", q[:m].tobytes(), z[:m].tobytes(), rs[:m].tobytes())
        rate = m / max(time.time() - t, 1e-6)
        m = int(min(q.shape[0], max(m, rate * 5.0)))
        t = time.time()
        okl = lc.verify_batch("k256", q[:m].tobytes(), z[:m].tobytes(), rs[:m].tobytes())
        dt2 = time.time() - t
        out["openssl"] = {"value": round(m / dt2, 1), "unit": "verifies/s", "cores": lc.threads_used(), "version": lc.lib().OpenSSL_version(0).decode()
                          if hasattr(lc.lib(), "OpenSSL_version") else "libcrypto", "sample": f"first {m} rows ({dt2:.1f} s), oracle/osslref.c (OpenMP)" if lc.driver() else f"first {m} rows ({dt2:.1f} s), ctypes + thread pool",
                          "matches_gpu_mask": bool(okl == exp[:m].tobytes())}
    except Exception as e:   # libcrypto missing on the box: the port figure stands alone
        out["openssl"] = {"unavailable": str(e)[:120]}
    return out


def other_configs(pkg, eng, dev, ts):
    """BASELINE configs 1, 2, 4, 5 at their own sizes, kernel-resident (CUDA events, 1 warm-up + 2 timed)."""
    import torch
    wl = pkg.workloads
    st = ts.cuda_stream
    res = []

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(reps):
            fn()
        e1.record(ts)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # config 1: k256 fixed-base, 2^16 scalars, constant-time path, compressed SEC1
    n = 1 << 16
    k = torch.from_numpy(wl.random_scalars(n, 32, 0xB2000001)).to(dev)
    o = torch.empty(n * 33, dtype=torch.uint8, device=dev)
    ms = timed(lambda: eng.mul_gen_dev("k256", n, k, o, pkg.FLAG_CT, st), 5)
    res.append({"config": "1: k256 G*k, 2^16 scalars, CT path, 33-B SEC1", "value": round(n / ms * 1e3, 1), "unit": "scalar-mul/s", "ms": round(ms, 4),
                "roofline_frac": roofline("mul_gen", "k256", n / ms * 1e3, ms, n, None)["frac"]})
    # config 2: k256 variable-base + batch_normalize, 2^20 projective inputs with random Z
    n = 1 << 20
    be = wl.EngineBackend(eng, "k256")
    xyz, kk = wl.make_mul_var_batch(be, "k256", n, 0xB2000002, projective=True)
    d_p, d_k = torch.from_numpy(xyz).to(dev), torch.from_numpy(kk).to(dev)
    o = torch.empty(n * 33, dtype=torch.uint8, device=dev)
    for name, fl in (("CT", pkg.FLAG_CT), ("VARTIME", 0)):
        ms = timed(lambda: eng.mul_var_dev("k256", n, d_p, None, d_k, o, None, fl | pkg.FLAG_PROJ, st))
        res.append({"config": f"2: k256 P*k + batch_normalize, 2^20 (X:Y:Z) inputs, {name}", "value": round(n / ms * 1e3, 1), "unit": "scalar-mul/s",
                    "ms": round(ms, 4), "roofline_frac": roofline("mul_var", "k256", n / ms * 1e3, ms, n, None)["frac"]})
    # config 4: p256 verify, 2^22 rows on this GPU (BASELINE quotes the same batch spread over 8 GPUs)
    n = 1 << 22
    q, z, rs, exp = wl.make_verify_batch(wl.EngineBackend(eng, "p256"), "p256", n, 0xB2000004)
    d_q, d_z, d_rs = torch.from_numpy(q).to(dev), torch.from_numpy(z).to(dev), torch.from_numpy(rs).to(dev)
    d_ok = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = timed(lambda: eng.ecdsa_verify_dev("p256", n, d_q, d_z, d_rs, d_ok, st))
    res.append({"config": "4: p256 ECDSA verify_prehash, 2^22 rows on 1 GPU", "value": round(n / ms * 1e3, 1), "unit": "verifies/s", "ms": round(ms, 4),
                "mask_ok": bool(np.array_equal(d_ok.cpu().numpy(), exp)),
                "roofline_frac": roofline("verify", "p256", n / ms * 1e3, ms, n, None)["frac"]})
    # config 5: p384 and sm2 variable-base, 2^20 each on this GPU (BASELINE quotes the same batches on 8 GPUs)
    for cname in ("p384", "sm2"):
        n = 1 << 20
        fb = 48 if cname == "p384" else 32
        pts, kk = wl.make_mul_var_batch(wl.EngineBackend(eng, cname), cname, n, 0xB2000005)
        d_p, d_k = torch.from_numpy(pts).to(dev), torch.from_numpy(kk).to(dev)
        o = torch.empty(n * (1 + 2 * fb), dtype=torch.uint8, device=dev)
        ms = timed(lambda: eng.mul_var_dev(cname, n, d_p, None, d_k, o, None, 0, st))
        res.append({"config": f"5: {cname} P*k, 2^20 on 1 GPU, uncompressed SEC1", "value": round(n / ms * 1e3, 1), "unit": "scalar-mul/s",
                    "ms": round(ms, 4), "roofline_frac": roofline("mul_var", cname, n / ms * 1e3, ms, n, None)["frac"]})
    # SURVEY §8 f4 tail: the primeorder template on 24-byte fields (P-192), verify at 2^20
    n = 1 << 20
    q, z, rs, exp = wl.make_verify_batch(wl.EngineBackend(eng, "p192"), "p192", n, 0xB2000008)
    d_q, d_z, d_rs = torch.from_numpy(q).to(dev), torch.from_numpy(z).to(dev), torch.from_numpy(rs).to(dev)
    d_ok = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = timed(lambda: eng.ecdsa_verify_dev("p192", n, d_q, d_z, d_rs, d_ok, st))
    res.append({"config": "f4: p192 ECDSA verify_prehash, 2^20 rows", "value": round(n / ms * 1e3, 1), "unit": "verifies/s", "ms": round(ms, 4),
                "mask_ok": bool(np.array_equal(d_ok.cpu().numpy(), exp))})
    # SURVEY §8f rows at 2^20 (k256): compressed-key verify, recovery, BIP340 (inputs made by the engine's own signer)
    n = 1 << 20
    rng = np.random.default_rng(0xB2000007)

    def scal():
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 0] &= 0x7F
        a[:, 31] |= 1
        return torch.from_numpy(a).to(dev)

    d_d, d_k, d_z = scal(), scal(), torch.from_numpy(rng.integers(0, 256, size=(n, 32), dtype=np.uint8)).to(dev)
    d_rs = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    d_id = torch.empty(n, dtype=torch.uint8, device=dev)
    d_ok = torch.empty(n, dtype=torch.uint8, device=dev)
    d_pub = torch.empty(n * 33, dtype=torch.uint8, device=dev)
    ms = timed(lambda: eng.ecdsa_sign_dev("k256", n, d_d, d_k, d_z, d_rs, d_id, d_ok, st))
    all_ok = bool(d_ok.all().item())
    res.append({"config": "f4: k256 ECDSA sign (CT fixed-base k*G + batched k^-1), 2^20", "value": round(n / ms * 1e3, 1), "unit": "signatures/s",
                "ms": round(ms, 4), "all_ok": all_ok})
    eng.mul_gen_dev("k256", n, d_d, d_pub, pkg.FLAG_CT, st)
    ms = timed(lambda: eng.ecdsa_verify_sec1_dev("k256", n, d_pub, 33, d_z, d_rs, d_ok, st))
    res.append({"config": "f1: k256 verify_prehash with compressed SEC1 keys (on-device decompression), 2^20", "value": round(n / ms * 1e3, 1),
                "unit": "verifies/s", "ms": round(ms, 4), "all_ok": bool(d_ok.all().item())})
    d_keys = torch.empty(n * 33, dtype=torch.uint8, device=dev)
    ms = timed(lambda: eng.ecdsa_recover_dev("k256", n, d_z, d_rs, d_id, d_keys, d_ok, 0, st))
    res.append({"config": "f2: k256 recover_from_prehash, 2^20", "value": round(n / ms * 1e3, 1), "unit": "recoveries/s", "ms": round(ms, 4),
                "all_ok": bool(d_ok.all().item()) and bool(torch.equal(d_keys, d_pub))})
    # BIP340: random (invalid) signatures over valid x-only keys exercise the full arithmetic path
    d_pkx = d_pub.view(n, 33)[:, 1:].contiguous()
    ms = timed(lambda: eng.schnorr_verify_dev(n, d_pkx, d_z, d_rs, d_ok, st))
    res.append({"config": "f2: k256 BIP340 Schnorr verify (random signatures over valid keys), 2^20", "value": round(n / ms * 1e3, 1),
                "unit": "verifies/s", "ms": round(ms, 4)})
    return res


if __name__ == "__main__":
    main()
