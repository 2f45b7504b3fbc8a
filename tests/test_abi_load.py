"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/ecb200.h
declares; the header and the Python mirror agree; no compute calls are made (no GPU here)."""
import ctypes
import os
import re

import pytest

import ecb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ecb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ecb200_[a-z0-9_]+)\s*\(", text)))


def test_header_matches_python_mirror():
    assert header_symbols() == sorted(ecb200.ABI_SYMBOLS)


def test_library_exports_every_symbol():
    if not os.path.exists(ecb200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(ecb200.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_size_helpers_need_no_gpu():
    lib = ecb200.load_library()
    assert [lib.ecb200_field_bytes(c) for c in range(7)] == [32, 32, 48, 32, 24, 28, 0]
    assert lib.ecb200_point_slot_bytes(0, 0) == 33 and lib.ecb200_point_slot_bytes(1, 0) == 65
    assert lib.ecb200_point_slot_bytes(2, 0) == 97 and lib.ecb200_point_slot_bytes(2, ecb200.FLAG_COMPRESSED) == 49
    assert ecb200.slot_bytes("k256") == 33 and ecb200.slot_bytes("sm2") == 65
    assert lib.ecb200_point_slot_bytes(4, 0) == 49 and ecb200.slot_bytes("p192", ecb200.FLAG_COMPRESSED) == 25
    assert lib.ecb200_version().startswith(b"ecb200")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ecb200.Ecb200Error):
        ecb200.Engine(0)


def test_bits2field():
    assert ecb200.bits2field("p256", b"\x01" * 48) == b"\x01" * 32
    assert ecb200.bits2field("p384", b"\x01" * 32) == b"\x00" * 16 + b"\x01" * 32
    with pytest.raises(ValueError):
        ecb200.bits2field("k256", b"\x01" * 15)


def test_shard_range():
    n = 1000003
    for w in (1, 2, 4, 8):
        rs = [ecb200.shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
