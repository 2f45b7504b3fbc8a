"""CPU tier: properties of the synthetic workload generator bench.py and the full-size tests rely on, exercised with the
oracle-backed stand-in engine (tests/fake_engine.py) at small sizes."""
import importlib

import numpy as np
import pytest

from oracle import ecoracle as o
from tests import fake_engine

wl = importlib.import_module("rustcrypto-elliptic-curves_b200.workloads")


@pytest.mark.parametrize("cname", ["k256", "p256"])
def test_verify_batch_is_prefix_stable_and_mask_matches_oracle(cname):
    """The rows of an n-row batch are the first n rows of a larger batch with the same seed (bench.py --impl reference and
    cpu_baseline time a PREFIX of the batch the GPU arm verifies), and the constructed mask is what the oracle computes."""
    be = wl.EngineBackend(fake_engine.OracleEngine(), cname)
    small = wl.make_verify_batch(be, cname, 48, 0xB2000003)
    big = wl.make_verify_batch(be, cname, 160, 0xB2000003)
    for a, b in zip(small, big):
        assert np.array_equal(a, b[:48])
    c = o.curve(cname)
    q, z, rs, exp = big
    assert o.batch_verify(c, q.tobytes(), z.tobytes(), rs.tobytes()) == exp.tobytes()
    assert 0 < int(exp.sum()) < 160
    # every corruption kind occurs in 160 rows (rows 5, 21, 37, 53, 69 are kinds 0..4)
    assert list(exp[[5, 21, 37]]) == [0, 0, 0] and exp[53] == (0 if cname == "k256" else 1) and exp[69] == 0


def test_random_scalars_prefix_stable():
    a, b = wl.random_scalars(100, 32, 7), wl.random_scalars(4096, 32, 7)
    assert np.array_equal(a, b[:100])


def test_mul_var_batch_projective_consistent():
    eng = fake_engine.OracleEngine()
    be = wl.EngineBackend(eng, "k256")
    xy, k = wl.make_mul_var_batch(be, "k256", 20, 5, projective=False)
    xyz, k2 = wl.make_mul_var_batch(be, "k256", 20, 5, projective=True)
    assert np.array_equal(k, k2)
    got, inf = eng.batch_normalize("k256", xyz)
    assert got == xy.tobytes() and not any(inf)
