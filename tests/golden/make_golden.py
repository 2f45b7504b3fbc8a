#!/usr/bin/env python3
"""Scrape the reference's own golden vectors for the hot path into JSON fixtures.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
Writes tests/golden/{group,field,ecdsa,wycheproof,misc}.json.  The GPU box has no /root/reference;
tests only read the committed JSON.  Nothing is copied but test DATA (hex vectors), with the source
file:line recorded next to each block.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
HEX = re.compile(r'hex!\(\s*((?:"[0-9a-fA-F\s]*"\s*)+)\)')


def hexes(text):
    out = []
    text = re.sub(r"//[^\n]*", "", text)   # line comments may sit between the string pieces
    for m in HEX.finditer(text):
        out.append("".join(re.findall(r'"([0-9a-fA-F\s]*)"', m.group(1))).replace(" ", "").replace("\n", "").lower())
    return out


def read(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def const_block(text, name):
    """Text of `pub const NAME ... = &[ ... ];`"""
    i = text.index("const " + name)
    j = text.index("\n];", i)
    return text[i:j]


def group_vectors(crate):
    t = read(f"{crate}/src/test_vectors/group.rs")
    add = hexes(const_block(t, "ADD_TEST_VECTORS"))
    mul = hexes(const_block(t, "MUL_TEST_VECTORS"))
    assert len(add) % 2 == 0 and len(mul) % 3 == 0
    return {
        "source": f"{crate}/src/test_vectors/group.rs",
        "add": [[add[i], add[i + 1]] for i in range(0, len(add), 2)],          # (i+1)*G = (x, y)
        "mul": [[mul[i], mul[i + 1], mul[i + 2]] for i in range(0, len(mul), 3)],  # k, x, y
    }


def field_vectors(crate):
    t = read(f"{crate}/src/test_vectors/field.rs")
    return {"source": f"{crate}/src/test_vectors/field.rs", "dbl": hexes(const_block(t, "DBL_TEST_VECTORS"))}


def ecdsa_vectors(crate):
    t = read(f"{crate}/src/test_vectors/ecdsa.rs")
    vecs = []
    for m in re.finditer(r"TestVector\s*\{(.*?)\n\s*\}", t, re.S):
        body = m.group(1)
        v = {}
        for fm in re.finditer(r"(\w+):\s*&(hex!\(.*?\))\s*,", body, re.S):
            v[fm.group(1)] = hexes(fm.group(2))[0]
        vecs.append(v)
    return {"source": f"{crate}/src/test_vectors/ecdsa.rs", "vectors": vecs}


def blobby_rows(path):
    """blobby 0.3 reader: VLQ ints, dedup table, 5 blobs per row."""
    with open(os.path.join(REF, path), "rb") as f:
        data = f.read()
    pos = 0

    def vlq():
        nonlocal pos
        b = data[pos]
        pos += 1
        val = b & 0x7F
        while b & 0x80:
            b = data[pos]
            pos += 1
            val = ((val + 1) << 7) + (b & 0x7F)
        return val

    d = vlq()
    dedup = []
    for _ in range(d):
        ln = vlq()
        dedup.append(data[pos:pos + ln])
        pos += ln
    blobs = []
    while pos < len(data):
        v = vlq()
        if v & 1:
            blobs.append(dedup[v >> 1])
        else:
            ln = v >> 1
            blobs.append(data[pos:pos + ln])
            pos += ln
    assert len(blobs) % 5 == 0, len(blobs)
    rows = []
    for i in range(0, len(blobs), 5):
        wx, wy, msg, sig, st = blobs[i:i + 5]
        assert len(st) == 1 and st[0] in (0, 1)
        rows.append([wx.hex(), wy.hex(), msg.hex(), sig.hex(), st[0]])
    return rows


def keccak256(data: bytes) -> bytes:
    """Keccak-256 (original padding 0x01, as sha3::Keccak256) — only to turn the reference's Ethereum-style
    vectors into prehashes; hashing is outside the hot path."""
    RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
          0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
          0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
          0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
    ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
    M = (1 << 64) - 1
    rol = lambda v, n: ((v << n) | (v >> (64 - n))) & M if n else v
    rate = 136
    msg = bytearray(data) + b"\x01" + b"\x00" * ((-len(data) - 2) % rate) + b"\x80" if (len(data) + 1) % rate else bytearray(data) + b"\x81"
    A = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i:off + 8 * i + 8], "little")
        for rc in RC:
            C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
            D = [C[(x - 1) % 5] ^ rol(C[(x + 1) % 5], 1) for x in range(5)]
            A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
            B = [[0] * 5 for _ in range(5)]
            for x in range(5):
                for y in range(5):
                    B[y][(2 * x + 3 * y) % 5] = rol(A[x][y], ROT[x][y])
            A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
            A[0][0] ^= rc
    return b"".join(A[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


def main():
    group = {c: group_vectors(c) for c in ("k256", "p256", "p384", "p192", "p224")}
    field = {c: field_vectors(c) for c in ("k256", "p256")}
    # risc0 8x32 KATs: k256/src/arithmetic/field/field_8x32_risc0.rs:225-303
    t = read("k256/src/arithmetic/field/field_8x32_risc0.rs")
    tt = t[t.index("mod tests"):]
    hx = hexes(tt)
    names = ["a", "b", "add", "add_negated", "negate", "mul", "square"]
    assert len(hx) == len(names), len(hx)
    field["k256_risc0_8x32"] = {"source": "k256/src/arithmetic/field/field_8x32_risc0.rs:225-303",
                                **dict(zip(names, hx))}
    ecdsa = {c: ecdsa_vectors(c) for c in ("k256", "p256", "p384", "p192", "p224")}
    wyche = {c: {"source": f"{c}/src/test_vectors/data/wycheproof.blb",
                 "hash": "sha384" if c == "p384" else "sha224" if c == "p224" else "sha256",
                 "rows": blobby_rows(f"{c}/src/test_vectors/data/wycheproof.blb")}
             for c in ("k256", "p256", "p384", "p224")}

    misc = {}
    # p256 prehash-longer-than-field accept vector: p256/src/ecdsa.rs:137-169
    t = read("p256/src/ecdsa.rs")
    blk = t[t.index("fn prehash_signer_verification_with_sha384"):t.index("fn scalar_blinding")]
    qx, qy, r, s, pre = hexes(blk)
    misc["p256_prehash_sha384_verify"] = {"source": "p256/src/ecdsa.rs:137-169", "qx": qx, "qy": qy,
                                          "r": r, "s": s, "prehash": pre, "expect": True}
    # RFC 6979 A.2.5 signatures (sign side; we verify them): p256/src/ecdsa.rs:98-118
    blk = t[t.index("fn rfc6979"):t.index("fn prehash_signer_signing_with_sha384")]
    d, sig_sample, sig_test = hexes(blk)
    misc["p256_rfc6979"] = {"source": "p256/src/ecdsa.rs:98-118", "d": d,
                            "sigs": [["sample", sig_sample], ["test", sig_test]], "hash": "sha256"}
    # p384 prehash-shorter-than-field case
    t = read("p384/src/ecdsa.rs")
    if "prehash_signer_verification_with_sha256" in t:
        i = t.index("fn prehash_signer_verification_with_sha256")
        j = t.index("\n    }\n", i)
        hx = hexes(t[i:j])
        qx, qy, r, s, pre = hx
        misc["p384_prehash_sha256_verify"] = {"source": "p384/src/ecdsa.rs:132-154", "qx": qx, "qy": qy,
                                              "r": r, "s": s, "prehash": pre, "expect": True}
    # k256 SEC1 fixtures: k256/src/arithmetic/affine.rs:374-379
    t = read("k256/src/arithmetic/affine.rs")
    i = t.index("mod tests")
    hx = hexes(t[i:])
    misc["k256_sec1_generator"] = {"source": "k256/src/arithmetic/affine.rs:374-379", "hexes": hx[:2]}
    # p256 SEC1 fixtures: p256/tests/affine.rs:12-28
    t = read("p256/tests/affine.rs")
    misc["p256_sec1_generator"] = {"source": "p256/tests/affine.rs:12-28", "hexes": hexes(t)[:2]}
    # sm2: d -> Q (sm2/tests/pkcs8.rs:43,48) and SM2DSA vector (sm2/tests/sm2dsa.rs:16-32)
    t = read("sm2/tests/pkcs8.rs")
    hx = hexes(t)
    misc["sm2_pkcs8"] = {"source": "sm2/tests/pkcs8.rs:43,48", "sec1_public": hx[0],
                         "d": [h for h in hx if len(h) == 64][0]}
    t = read("sm2/tests/sm2dsa.rs")
    hx = hexes(t)
    misc["sm2dsa"] = {"source": "sm2/tests/sm2dsa.rs:16-32", "sec1_public": hx[0], "sig": hx[1],
                      "identity": "example@rustcrypto.org", "msg": "testing", "expect": True}
    # k256 bench fixed scalars: k256/benches/scalar.rs:14-35, ecdsa.rs:13-37
    misc["k256_bench_scalars"] = {"source": "k256/benches/scalar.rs,ecdsa.rs",
                                  "hexes": hexes(read("k256/benches/scalar.rs")) + hexes(read("k256/benches/ecdsa.rs"))}

    # ---- "next" rows: BIP340 vectors, public-key recovery vectors
    t = read("k256/src/schnorr.rs")
    tt = t[t.index("const BIP340_SIGN_VECTORS"):t.index("fn bip340_sign_vectors")]
    sign = []
    for blk in tt.split("SignVector {")[1:]:
        idx = int(re.search(r"index:\s*(\d+)", blk).group(1))
        sk, pk, aux, msg, sig = hexes(blk)
        sign.append({"index": idx, "secret_key": sk, "public_key": pk, "aux_rand": aux, "message": msg, "signature": sig})
    tt = t[t.index("const BIP340_VERIFY_VECTORS"):t.index("fn bip340_verify_vectors")]
    ver = []
    for blk in tt.split("VerifyVector {")[1:]:
        idx = int(re.search(r"index:\s*(\d+)", blk).group(1))
        pk, msg, sig = hexes(blk)
        ver.append({"index": idx, "public_key": pk, "message": msg, "signature": sig,
                    "valid": re.search(r"valid:\s*(true|false)", blk).group(1) == "true"})
    t = read("k256/src/ecdsa.rs")
    tt = t[t.index("const RECOVERY_TEST_VECTORS"):t.index("fn public_key_recovery")]
    rec = []
    for blk in tt.split("RecoveryTestVector {")[1:]:
        pk, sig = hexes(blk)
        msg = re.search(r'msg:\s*b"([^"]*)"', blk).group(1)
        yo, xr = re.search(r"RecoveryId::new\((true|false),\s*(true|false)\)", blk).groups()
        rec.append({"pk": pk, "msg": msg, "hash": "sha256", "sig": sig, "recid": (yo == "true") | ((xr == "true") << 1)})
    # module-doc example (k256/src/ecdsa.rs:113-140): Keccak256 prehash, recovery id 1
    doc = t[t.index("### Recovering a [`VerifyingKey`] from a signature"):t.index("assert_eq!(recovered_key, expected_key)")]
    dh = hexes(re.sub(r"(?m)^//!", "", doc))
    rec.append({"pk": dh[1], "msg": "example message", "hash": "keccak256", "sig": dh[0], "recid": 1})
    tt = t[t.index("fn ethereum_end_to_end_example"):t.index("mod wycheproof")]
    eh = hexes(tt)
    eth = {"source": "k256/src/ecdsa.rs:310-340", "d": eh[0], "msg": eh[1], "sig": eh[2], "recid": 0, "hash": "keccak256",
           "prehash": keccak256(bytes.fromhex(eh[1])).hex()}
    for r_ in rec:
        m_ = r_["msg"].encode()
        r_["prehash"] = (keccak256(m_) if r_["hash"] == "keccak256" else __import__("hashlib").sha256(m_).digest()).hex()
    nxt = {"bip340_sign": {"source": "k256/src/schnorr.rs:217-289", "vectors": sign},
           "bip340_verify": {"source": "k256/src/schnorr.rs:308-449", "vectors": ver},
           "k256_recovery": {"source": "k256/src/ecdsa.rs:113-140,278-310", "vectors": rec},
           "k256_ethereum_sign_recover": eth}
    with open(os.path.join(OUT, "next.json"), "w") as f:
        json.dump(nxt, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("next: bip340 sign", len(sign), "verify", len(ver), "recovery", len(rec))

    for name, obj in (("group", group), ("field", field), ("ecdsa", ecdsa), ("wycheproof", wyche), ("misc", misc)):
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(obj, f, indent=0, separators=(",", ":"))
            f.write("\n")
        print(name, os.path.getsize(os.path.join(OUT, name + ".json")), "bytes")
    for c in wyche:
        rows = wyche[c]["rows"]
        print(c, "wycheproof rows", len(rows), "pass", sum(r[4] for r in rows))
    for c in group:
        print(c, "add", len(group[c]["add"]), "mul", len(group[c]["mul"]))
    for c in ecdsa:
        print(c, "ecdsa", len(ecdsa[c]["vectors"]))


if __name__ == "__main__":
    main()
