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


def main():
    group = {c: group_vectors(c) for c in ("k256", "p256", "p384")}
    field = {c: field_vectors(c) for c in ("k256", "p256")}
    # risc0 8x32 KATs: k256/src/arithmetic/field/field_8x32_risc0.rs:225-303
    t = read("k256/src/arithmetic/field/field_8x32_risc0.rs")
    tt = t[t.index("mod tests"):]
    hx = hexes(tt)
    names = ["a", "b", "add", "add_negated", "negate", "mul", "square"]
    assert len(hx) == len(names), len(hx)
    field["k256_risc0_8x32"] = {"source": "k256/src/arithmetic/field/field_8x32_risc0.rs:225-303",
                                **dict(zip(names, hx))}
    ecdsa = {c: ecdsa_vectors(c) for c in ("k256", "p256", "p384")}
    wyche = {c: {"source": f"{c}/src/test_vectors/data/wycheproof.blb",
                 "hash": "sha384" if c == "p384" else "sha256",
                 "rows": blobby_rows(f"{c}/src/test_vectors/data/wycheproof.blb")}
             for c in ("k256", "p256", "p384")}

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
