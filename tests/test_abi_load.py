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


def test_header_constants_match_python_and_cpp_mirrors():
    """Curve ids, flags and decode modes: include/ecb200.h is the source; the ctypes layer and the C++ mirror must agree."""
    text = open(os.path.join(ROOT, "include", "ecb200.h")).read()
    enum = {k: int(v) for k, v in re.findall(r"ECB200_(K256|P256|P384|SM2|P192|P224) = (\d+)", text)}
    assert enum == {"K256": ecb200.K256, "P256": ecb200.P256, "P384": ecb200.P384, "SM2": ecb200.SM2, "P192": ecb200.P192, "P224": ecb200.P224}
    assert sorted(enum.values()) == list(range(6))
    flags = {k: int(v) for k, v in re.findall(r"#define ECB200_FLAG_(\w+) (\d+)u", text)}
    assert flags == {"CT": ecb200.FLAG_CT, "COMPRESSED": ecb200.FLAG_COMPRESSED, "UNCOMPRESSED": ecb200.FLAG_UNCOMPRESSED, "PROJ": ecb200.FLAG_PROJ}
    hpp = open(os.path.join(ROOT, "include", "ecb200.hpp")).read()
    for name, fb in (("K256", 32), ("P256", 32), ("P384", 48), ("SM2", 32), ("P192", 24), ("P224", 28)):
        m = re.search(r"static constexpr int ID = ECB200_%s;\s*static constexpr size_t FB = (\d+);" % name, hpp)
        assert m and int(m.group(1)) == fb == ecb200.field_bytes(enum[name]), name


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
