"""Oracle evaluation of a sample of rows on every host core (test infrastructure; the pure-Python oracle does one
scalar multiplication in 2.4 - 7.3 ms, so a 2^16-row sample of a BASELINE-size batch needs the whole host).

Workers are spawned (not forked): the parent process holds a CUDA context."""
import multiprocessing as mp
import os


def _mul_var(args):
    cname, proj, pts, ks = args
    from oracle import ecoracle as o
    c = o.curve(cname)
    return o.batch_mul_var_proj(c, pts, ks) if proj else o.batch_mul_var_affine(c, pts, None, ks)


def _mul_gen(args):
    cname, ks, compress = args
    from oracle import ecoracle as o
    return o.batch_mul_gen(o.curve(cname), ks, compress)


def _split(n, parts):
    step = max(1, (n + parts - 1) // parts)
    return [(lo, min(n, lo + step)) for lo in range(0, n, step)]


def _run(fn, jobs, procs):
    procs = procs or os.cpu_count() or 1
    if procs == 1 or len(jobs) == 1:
        return b"".join(fn(j) for j in jobs)
    with mp.get_context("spawn").Pool(min(procs, len(jobs))) as pool:
        return b"".join(pool.map(fn, jobs))


def mul_var(cname, fb, pts: bytes, ks: bytes, proj: bool, procs=None) -> bytes:
    """SEC1 slots (curve default encoding) of k_i * P_i for the given rows, computed by oracle/ecoracle.py."""
    n = len(ks) // fb
    pb = fb * (3 if proj else 2)
    jobs = [(cname, proj, pts[lo * pb:hi * pb], ks[lo * fb:hi * fb]) for lo, hi in _split(n, 4 * (procs or os.cpu_count() or 1))]
    return _run(_mul_var, jobs, procs)


def mul_gen(cname, fb, ks: bytes, compress=None, procs=None) -> bytes:
    n = len(ks) // fb
    jobs = [(cname, ks[lo * fb:hi * fb], compress) for lo, hi in _split(n, 4 * (procs or os.cpu_count() or 1))]
    return _run(_mul_gen, jobs, procs)
