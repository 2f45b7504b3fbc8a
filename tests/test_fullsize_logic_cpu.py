"""CPU tier: the Python around the BASELINE-size GPU tests (edge rows, libcrypto / oracle comparisons, samplers) run at a
tiny size against the oracle-backed stand-in engine, so that a bug in the checkers is found here and not on the GPU box.
Nothing is claimed about the CUDA path (tests/fake_engine.py answers from oracle/ecoracle.py)."""
import pytest

from tests import fake_engine

pytest.importorskip("oracle.libcrypto_ref")


@pytest.fixture()
def full(monkeypatch):
    import tests.test_gpu_fullsize as m
    monkeypatch.setattr(m, "LOG2", 6)
    monkeypatch.setattr(m, "SAMPLE_LOG2", 5)
    return m


def test_config2_logic(full):
    full.test_config2_k256_mul_var_fullsize(fake_engine.OracleEngine())


@pytest.mark.parametrize("cname,seed", [("p384", 0xB2000005), ("sm2", 0xB2000006)])
def test_config5_logic(full, cname, seed):
    full.test_config5_primeorder_mul_var_fullsize(fake_engine.OracleEngine(), cname, seed)
