"""bench.py's contract, as far as it can be checked without a GPU: the reference arm's JSON line (it runs on the host cores),
its behaviour on the other ranks of a torchrun launch, the refusal of our arm to run without a CUDA device (no CPU fallback),
and the roofline block computed from the tracked ncu summaries (profiles/summary.json) for every BASELINE configuration."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _bench_module():
    spec = importlib.util.spec_from_file_location("ecb200_bench_py", BENCH)     # by path: `bench/` (the peak micro-benchmark) shares the name
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(args, env_extra=None, timeout=280):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_line_carries_the_contract_keys():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "rank 0 prints ONE JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "verifies/s" and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "configs[2]" in d["config"]["workload"] and "bounded_sample" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"] and cb["unit"] == d["unit"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the arm would run")
    p = _run(["--steps", "1", "--warmup", "1", "--no-others", "--no-cpu"])
    assert p.returncode != 0, "no CPU fallback: the product arm must fail loudly without a device"
    assert p.stdout.strip() == "", "no result line may be printed"
    assert "no CPU fallback" in p.stderr or "CUDA" in p.stderr


@pytest.mark.parametrize("op,curve,ct,rowpath", [
    ("verify", "k256", False, False), ("verify", "k256", False, True), ("verify", "p256", False, False), ("verify", "p256", False, True),
    ("mul_gen", "k256", True, False), ("mul_var", "k256", False, False), ("mul_var", "k256", True, False),
    ("mul_var", "p384", False, False), ("mul_var", "sm2", False, False)])
def test_roofline_block_is_backed_by_a_tracked_ncu_summary(op, curve, ct, rowpath):
    """every BASELINE configuration has executed-instruction counts in profiles/summary.json whose source file is tracked, and the
    block bench.py derives from them is a hardware fraction (0 < frac < 1 at the recorded durations)"""
    b = _bench_module()
    summ = json.load(open(os.path.join(ROOT, "profiles", "summary.json")))
    key = b.prof_key(op, curve, ct, rowpath)
    assert key in summ, key
    prof = summ[key]
    assert prof.get("kernels") and prof.get("n_rows") and prof.get("wide_macs_per_row") > 0
    src = prof.get("source")
    assert src and os.path.isfile(os.path.join(ROOT, src)), "the ncu summary the counts come from must be a tracked file: %r" % src
    dom = max(prof["kernels"], key=lambda k: k["ms"])
    step_ms = sum(k["ms"] for k in prof["kernels"])
    r = b.roofline(op, curve, prof["n_rows"], step_ms, 1965.0, kernel_ms=dom["ms"], ct=ct, io=(5 * 32, 1), rowpath=rowpath)
    assert r["bound"] == "imad" and r["executed_source"] == src
    assert 0.05 < r["frac"] < 1.0 and 0.05 < r["step_executed_frac"] < 1.0 and r["frac_of_nominal"] < r["frac"]
    assert r["peak"] > 0 and r["peak_nominal"] > r["peak"] and r["traffic"] is not None
    assert r["hbm"]["frac"] < 0.05, "three orders of magnitude below the bandwidth roofline"
