#!/usr/bin/env python
"""Write the first 2^L rows of bench.py's verify batch (same seed, keys and corruption pattern: workloads.make_verify_batch is
prefix-stable) as raw files for rust/bench (the reference arm with the real crates, for a machine that has cargo):
    python tools/dump_batch.py k256 18 /tmp/batch      ->  q.bin z.bin rs.bin expected.bin
Runs without a GPU: k*G comes from the C++ port (test infrastructure), the mod-n algebra from Python integers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util

spec = importlib.util.spec_from_file_location("ecb200_bench_py", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def main():
    curve, lg, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    wl = bench.importlib_pkg().workloads
    seed = bench.CASES[("verify", curve)]["seed"]
    q, z, rs, exp = wl.make_verify_batch(bench.PyBackend(curve), curve, 1 << lg, seed)
    os.makedirs(out, exist_ok=True)
    for name, a in (("q.bin", q), ("z.bin", z), ("rs.bin", rs), ("expected.bin", exp)):
        a.tofile(os.path.join(out, name))
    print("%d rows of %s (seed %#x) -> %s" % (1 << lg, curve, seed, out))


if __name__ == "__main__":
    main()
