"""The C++ host mirror (include/ecb200.hpp) — the reference's trait surface over the C ABI, in the language class of
the reference (compiled code; no Rust toolchain in this image).  tests/cabi/host_mirror.cpp replays the reference's
own hot-path tests through it: group vectors, lincomb / mul_by_generator / batch_normalize consistency tests, FIPS and
Wycheproof verification, signing KATs, recovery, BIP340 and SM2DSA vectors.  The fixture it reads is written here from
tests/golden/*.json (scraped from the reference) with the oracle's expectations beside them."""
import hashlib
import os
import subprocess

import pytest

from oracle import ecoracle as o
from tests import nextrows as nr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_exe(tmp_path):
    import ecb200
    ecb200.load_library()                      # raises if libecb200.so has not been built
    exe = str(tmp_path / "host_mirror")
    libdir = os.path.dirname(ecb200.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cabi", "host_mirror.cpp"), "-o", exe,
                           "-L", libdir, "-lecb200", "-Wl,-rpath," + libdir])
    return exe


def hx(v, n):
    return int(v).to_bytes(n, "big").hex()


def write_fixture(path, golden):
    lines = []
    for cname in ("k256", "p256", "p384", "sm2", "p192", "p224"):
        c = o.curve(cname)
        fb = c.fb
        # group vectors (sm2: the reference has none; the oracle, pinned by libcrypto in the CPU tier, supplies them)
        if cname in golden["group"]:
            adds = [(int(x, 16), int(y, 16)) for x, y in golden["group"][cname]["add"]]
            muls = [(int(k, 16), int(x, 16), int(y, 16)) for k, x, y in golden["group"][cname]["mul"]]
        else:
            adds = [o.mul_gen(c, i + 1) for i in range(20)]
            muls = [(k,) + o.mul_gen(c, k) for k in (0x18EBBB95EED0E13, c.n - 1, c.n >> 1, (1 << 128) - 1)]
        for x, y in adds:
            lines.append("add %s %s %s" % (cname, hx(x, fb), hx(y, fb)))
        for k, x, y in muls:
            lines.append("mul %s %s %s %s" % (cname, hx(k, fb), hx(x, fb), hx(y, fb)))
        # SEC1 decoding
        slots, stride, status, xy = nr.decode_cases(c)
        for i in range(len(status)):
            s = slots[i * stride:(i + 1) * stride]
            ln = 1 if s[0] == 0 and not any(s) else 1 + 2 * fb if s[0] == 4 else 1 + fb
            lines.append("decode %s %s %02x %s %s" % (cname, s[:ln].hex(), status[i], xy[i * 2 * fb:i * 2 * fb + fb].hex(), xy[i * 2 * fb + fb:(i + 1) * 2 * fb].hex()))
        # verification: FIPS vectors + flipped-s negative (new_verification_test!), prehash-length cases, Wycheproof
        ver = []
        if cname in golden["ecdsa"]:
            for v in golden["ecdsa"][cname]["vectors"]:
                Q, m, r, s = (int(v["q_x"], 16), int(v["q_y"], 16)), bytes.fromhex(v["m"]), int(v["r"], 16), int(v["s"], 16)
                ver.append((Q, m, r, s))
                ver.append((Q, m, r, s ^ 1))
        if cname == "p256":
            m = golden["misc"]["p256_prehash_sha384_verify"]
            ver.append(((int(m["qx"], 16), int(m["qy"], 16)), bytes.fromhex(m["prehash"]), int(m["r"], 16), int(m["s"], 16)))
            ver.append((c.G, b"\x01" * 15, 1, 1))                                  # prehash too short -> Err
        if cname == "p384":
            m = golden["misc"]["p384_prehash_sha256_verify"]
            ver.append(((int(m["qx"], 16), int(m["qy"], 16)), bytes.fromhex(m["prehash"]), int(m["r"], 16), int(m["s"], 16)))
        if cname in golden["wycheproof"]:
            blob = golden["wycheproof"][cname]
            hf = getattr(hashlib, blob["hash"])
            for wx, wy, msg, sig, flag in blob["rows"]:
                rs = o.der_parse_strict(bytes.fromhex(sig), c)
                if rs is None or max(rs) >= 1 << (8 * fb):
                    assert not flag
                    continue
                r, s = rs
                Q = (int.from_bytes(bytes.fromhex(wx)[-fb:], "big"), int.from_bytes(bytes.fromhex(wy)[-fb:], "big"))
                digest = hf(bytes.fromhex(msg)).digest()
                if c.low_s and 1 <= s < c.n and s > c.n >> 1:
                    ver.append((Q, digest, r, s))
                    s = c.n - s                                                    # the reference's runner normalises (k256 ecdsa.rs:389)
                ver.append((Q, digest, r, s))
        if cname == "sm2":                                                         # sm2-as-ECDSA: oracle-signed rows
            import random
            rng = random.Random(44)
            for i in range(6):
                d, k, z, (r, s, _) = nr.make_sig(c, rng)
                ver.append((o.mul_gen(c, d), z, r, s))
                ver.append((o.mul_gen(c, d), z, r, s ^ 2))
        for Q, h, r, s in ver:
            exp = o.verify_prehash(c, Q, h, r, s)
            lines.append("verify %s %s %s %s %s %s %02x" % (cname, hx(Q[0], fb), hx(Q[1], fb), h.hex(), hx(r, fb), hx(s, fb), 1 if exp else 0))
        # signing KATs (ok rows only: d, k in [1, n-1])
        db, kb, zb, rs, rid, ok = nr.sign_cases(c, golden)
        for i in range(len(ok)):
            if ok[i]:
                lines.append("sign %s %s %s %s %s %s %02x" % (cname, db[i * fb:(i + 1) * fb].hex(), kb[i * fb:(i + 1) * fb].hex(), zb[i * fb:(i + 1) * fb].hex(),
                                                             rs[i * 2 * fb:i * 2 * fb + fb].hex(), rs[i * 2 * fb + fb:(i + 1) * 2 * fb].hex(), rid[i]))
    # recovery (k256 vectors of k256/src/ecdsa.rs:278-343 plus the oracle's failing rows)
    c = o.K256
    zb, rsb, ids, exp_keys, exp_ok = nr.recover_cases(c, golden)
    for i in range(len(ids)):
        want = exp_keys[i * 33:(i + 1) * 33].hex() if exp_ok[i] else "-"
        lines.append("recover k256 %s %s %s %02x %s" % (zb[i * 32:(i + 1) * 32].hex(), rsb[i * 64:i * 64 + 32].hex(), rsb[i * 64 + 32:(i + 1) * 64].hex(), ids[i], want))
    pkb, eb, sb, exp = nr.schnorr_cases(golden)
    for i in range(len(exp)):
        lines.append("schnorr k256 %s %s %s %02x" % (pkb[i * 32:(i + 1) * 32].hex(), eb[i * 32:(i + 1) * 32].hex(), sb[i * 64:(i + 1) * 64].hex(), exp[i]))
    qb, eb, rsb, exp = nr.sm2dsa_cases(golden)
    for i in range(len(exp)):
        lines.append("sm2dsa sm2 %s %s %s %s %s %02x" % (qb[i * 64:i * 64 + 32].hex(), qb[i * 64 + 32:(i + 1) * 64].hex(), eb[i * 32:(i + 1) * 32].hex(),
                                                        rsb[i * 64:i * 64 + 32].hex(), rsb[i * 64 + 32:(i + 1) * 64].hex(), exp[i]))
    with open(path, "w") as f:
        f.write("# kind curve fields... (written by tests/test_host_mirror.py)\n" + "\n".join(lines) + "\n")
    return len(lines)


def test_host_mirror_compiles_and_host_logic(tmp_path):
    """CPU tier: the header compiles warning-free against the C ABI, links with libecb200.so, its host-side logic
    (Scalar / Signature range rules, bits2field, SEC1 framing, shard_range) holds, and without a CUDA device the
    engine refuses to start (no CPU fallback)."""
    import torch
    exe = build_exe(tmp_path)
    args = [exe, "--host-only"] + ([] if torch.cuda.is_available() else ["--expect-no-device"])
    out = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 failures" in out.stdout
    if not torch.cuda.is_available():
        assert "engine refused" in out.stdout


def test_fixture_builder(tmp_path, golden):
    """CPU tier: the fixture covers every kind and curve the C++ program replays."""
    p = tmp_path / "fx.txt"
    n = write_fixture(str(p), golden)
    kinds = {}
    for line in p.read_text().splitlines()[1:]:
        k, c = line.split()[:2]
        kinds[(k, c)] = kinds.get((k, c), 0) + 1
    for c in ("k256", "p256", "p384", "sm2", "p192", "p224"):
        for k in ("add", "mul", "decode", "verify", "sign"):
            assert kinds.get((k, c), 0) >= 4, (k, c)
    assert kinds[("verify", "k256")] > 200 and kinds[("verify", "p256")] > 200 and kinds[("verify", "p384")] > 200
    assert kinds[("schnorr", "k256")] >= 15 and kinds[("sm2dsa", "sm2")] >= 10 and kinds[("recover", "k256")] >= 10
    assert n > 1000


@pytest.mark.gpu
def test_host_mirror_on_gpu(tmp_path, golden):
    """GPU tier: the reference's tests through the C++ mirror on the B200, no Python between the test and the C ABI."""
    exe = build_exe(tmp_path)
    fx = str(tmp_path / "fixture.txt")
    write_fixture(fx, golden)
    out = subprocess.run([exe, fx], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "host mirror ok" in out.stdout
    for c in ("k256", "p256", "p384", "sm2", "p192", "p224"):
        assert "%s: ok" % c in out.stdout
