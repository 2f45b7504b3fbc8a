#!/usr/bin/env python
"""Static side-channel audit of the SECRET-SCALAR kernels, on the SASS ptxas actually produced.

north_star: "the secret-scalar mul path is fixed-window with no secret-dependent branches or addresses" - the discipline of
k256/src/arithmetic/mul.rs:92-127 and primeorder/src/projective.rs:127-147.  The C++ source is written that way (masked
scans, masked selects), but nothing in C++ forces ptxas to keep a ternary a SEL; this tool looks at the machine code:

  * every CONDITIONAL control transfer (predicated BRA / EXIT / RET / CALL, BRA.U on a uniform predicate, any indirect
    BRX / JMX) of k_mul_var<*, true>, k_mul_gen_smem<*, true>, k_gen_half<*, true>, k_sum_normalize<*> and k_sign_finish<*> - including the device functions they
    call - is listed with the source line nvdisasm attributes it to (-lineinfo) and the instruction that set its predicate;
  * each must sit on a source line matching an ALLOW pattern: row guard (tid >= n), loops over public counters, public flags
    (F_PROJ, the `inf` byte), validity of PUBLIC inputs (point on curve / coordinates < p), the mbarrier wait of the TMA copy,
    and loop bounds of the Montgomery-trick bodies (rows per thread).  Anything else is a FINDING and the exit code is 1;
  * every PREDICATED memory access (@P LDL / LDS / LDG / STL ...) is treated the same way: ptxas predicated the table-scan loads
    of round 1 on the secret comparison (a masked select turned back into a conditional load), which no branch scan sees;
  * LDL / STL / LDS / LDG whose address register is produced by arithmetic on a value loaded from the secret buffers cannot
    be proven absent by pattern matching; the dynamic counterpart (scripts/ct_audit.sh: identical instruction, branch and
    memory-sector counts for different secrets under ncu) covers addresses - and is what found the predicated loads.

usage: ct_sass_audit.py [--out profiles/r02_ct_sass_audit.md]      (needs the built objects in rustcrypto-elliptic-curves_b200/_build)
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "rustcrypto-elliptic-curves_b200", "_build")
CURVES = ["k256", "p256", "p384", "sm2", "p192", "p224"]
KERNELS = [r"k_mul_varINS_\w+ELb1E", r"k_mul_gen_smemINS_\w+ELb1E", r"k_sign_finishINS_", r"k_gen_halfINS_\w+ELb1E", r"k_sum_normalizeINS_"]
# source-line patterns under which a conditional branch is legitimate in a secret-scalar kernel (all public quantities)
ALLOW = [
    (r"if \(tid >= n\) return", "row guard (public batch size)"),
    (r"if \(i >= n\) break", "row guard inside the rows-per-thread loop (public batch size)"),
    (r"if \(!live\) return", "row guard of the shuffle-fetch kernel (public batch size; rows past the end are clamped, computed and not stored)"),
    (r"\bfor \(", "loop over a public counter"),
    (r"flags & F_PROJ", "public flag: projective vs affine input"),
    (r"inf && inf\[tid\]", "public identity flag of the input point"),
    (r"if \(!ok\) G::set_identity\(p\)", "validity of the PUBLIC input point (on curve, coordinates < p)"),
    (r"while \(!done\)", "mbarrier wait of the TMA table copy"),
    (r"mbarrier|cp\.async\.bulk|smem_addr\(", "TMA table copy (inline PTX: mbarrier arrive / try_wait)"),
    (r"\(j & 1\) == 0", "parity of the table-construction counter"),
    (r"if \(i != 32\)|if \(w != top\)|if \(i == 8 \* L\)|if \(w == top\)", "first / last window of a fixed-length loop (public counter)"),
    (r"if \(nib\) mul\(acc, acc, tab\[nib\]\)", "nibble of the PUBLIC inversion exponent (n - 2 / p - 2)"),
    (r"if \(threadIdx\.x == 0\)|if \(\(int\)threadIdx\.x == chain_thread\)|if \(lane == 31\)", "thread-index test"),
    (r"INV::COOPERATIVE && cnt == 0", "rows-per-thread count (public batch size)"),
    (r"__launch_bounds__|__global__", "kernel prologue / epilogue"),
    (r"if \(n <= 0\)", "launcher"),
    (r"return on_curve\(a\) && ok|bool ok = F::from_limbs|ok = F::from_limbs|return F::eq\(lhs, rhs\)", "validation of the PUBLIC input point"),
    (r"if \(invalid\) invalid\[tid\]", "optional output pointer (public)"),
    # k_sum_normalize (tail of the split fixed-base path): output format switches, and the identity test of the RESULT point -
    # the result is what the call returns in clear (an all-zero SEC1 slot / the ok byte of a signature), so it is public by
    # construction; the complete addition and the inversion before it carry no branch at all
    (r"if \(mode == NORM_SEC1\)|mode == NORM_XY_BYTES|mode == NORM_AFF_STRIDED", "output format (public call argument)"),
    (r"if \(inf\) \{|if \(isinf\) \{|out_inf\[i\] = isinf", "identity test of the result point, which the call outputs in clear"),
    (r"out\[0\] = \(u8\)\(2u \+ \(t\[0\] & 1u\)\)|if \(compress\)", "SEC1 compression flag (public call argument)"),
    (r"if \(sum_with\)", "optional second operand array (public call shape)"),
    (r"e\.v\[i - 1\] = \(int\)ce & M30; ce >>= 30;|g\.v\[i - 1\] = \(int\)cg & M30; cg >>= 30;|g >>= 1; u <<= 1; v <<= 1;",
     "back edge of a fixed-trip-count limb / divstep loop of safegcd.cuh (public counter; ptxas keeps the 13-limb loop rolled)"),
]


ADDRESS_ALIGNMENT = r"reinterpret_cast<unsigned long long>\(p\) & A\) == 0ull"


def sh(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw)


def disasm(curve, tmp):
    obj = os.path.join(BUILD, "curve_%s.o" % curve)
    if not os.path.exists(obj):
        return None
    sh(["cuobjdump", "-xelf", "all", obj], cwd=tmp)
    cubin = os.path.join(tmp, "curve_%s.sm_100a.cubin" % curve)
    return sh(["nvdisasm", "-g", "-c", cubin]).stdout


_src_cache = {}


def src_line(path, line):
    if path not in _src_cache:
        try:
            _src_cache[path] = open(path).read().splitlines()
        except OSError:
            _src_cache[path] = []
    ls = _src_cache[path]
    return ls[line - 1].strip() if 0 < line <= len(ls) else ""


INS = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;")
LOC = re.compile(r'//## File "([^"]+)", line (\d+)(.*)')


def audit_kernel(name, body):
    """body: lines of one .text section.  Returns (n_instructions, [finding dicts], [allowed dicts])."""
    ins = []          # (addr, text, (file, line), inlined chain text)
    cur = ("", 0)
    chain = ""
    for l in body:
        m = LOC.search(l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            chain = m.group(3)
            continue
        m = INS.match(l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), cur, chain))
    cond = []
    for i, (addr, text, loc, ch) in enumerate(ins):
        toks = text.split()
        guard = toks[0] if toks[0].startswith("@") else None
        op = toks[1] if guard else toks[0]
        base = op.split(".")[0]
        is_mem = base in ("LDL", "LDS", "LDG", "LD", "STL", "STS", "STG", "ST", "LDSM", "ATOMS", "ATOMG", "RED")
        if base not in ("BRA", "EXIT", "RET", "CALL", "BRX", "JMX", "JMP", "BREAK") and not (is_mem and guard and guard != "@PT"):
            continue
        pred = None
        if base in ("BRX", "JMX"):
            pred = "indirect"
        elif guard and guard not in ("@PT",):
            pred = guard.lstrip("@!")
        elif base == "BRA":
            rest = " ".join(toks[(2 if guard else 1):])
            m = re.match(r"(!?U?P\d+),", rest)
            if m:
                pred = m.group(1).lstrip("!")
        if pred is None:
            continue
        # nearest earlier instruction that writes the predicate (linear scan: a heuristic, good enough to show the comparison)
        setter = ""
        if pred != "indirect":
            for j in range(i - 1, max(-1, i - 400), -1):
                t = ins[j][1]
                tk = t.split()
                body_ = tk[1:] if not tk[0].startswith("@") else tk[2:]
                if any(re.fullmatch(re.escape(pred) + r",?", x) for x in body_[:2]):
                    setter = "%04x: %s   [%s:%d]" % (ins[j][0], t, os.path.basename(ins[j][2][0]), ins[j][2][1])
                    break
        cond.append({"addr": addr, "text": text, "file": loc[0], "line": loc[1], "src": src_line(loc[0], loc[1]), "setter": setter, "chain": ch, "mem": is_mem})
    findings, allowed = [], []
    for c in cond:
        why = next((w for pat, w in ALLOW if re.search(pat, c["src"])), None)
        if why is None and c["mem"] and re.search(r"\bUP\d", c["setter"].split("[")[0]):
            why = "memory access predicated on a UNIFORM predicate (kernel parameters / warp-uniform counters)"
        if why is None and (c["text"].startswith("BRA.U") or c["mem"]) and c["setter"]:
            # a branch on a UNIFORM predicate (kernel parameters, CTA index, warp-uniform counters): judged by the line that
            # computed the predicate, since nvdisasm attributes the branch itself to the statement that follows it
            m = re.search(r"\[([\w.]+):(\d+)\]$", c["setter"])
            if m:
                path = next((p for p in _src_cache if os.path.basename(p) == m.group(1)), None)
                if path:
                    sl = src_line(path, int(m.group(2)))
                    why = next((w + " (uniform predicate, set at %s:%s)" % (m.group(1), m.group(2)) for pat, w in ALLOW if re.search(pat, sl)), None)
        if why is None and c["setter"]:
            # the alignment dispatch of load_be / store_be: the predicate is an AND of the element address (buffer base + row
            # index, public) with a constant; nvdisasm attributes the branch to the closing brace of the if, so it is judged
            # by the line that set the predicate
            m = re.search(r"\[([\w.]+):(\d+)\]$", c["setter"])
            path = m and next((p for p in _src_cache if os.path.basename(p) == m.group(1)), None)
            if path and re.search(ADDRESS_ALIGNMENT, src_line(path, int(m.group(2)))) and re.search(r"LOP3\.LUT P\d, RZ, R\d+, 0x[37f],", c["setter"]):
                why = "alignment of a buffer ADDRESS (base pointer + row index, public): word or byte path of load_be / store_be (set at %s:%s)" % (m.group(1), m.group(2))
        if why is None and "SYNCS.PHASECHK" in c["setter"]:
            why = "mbarrier try_wait of the TMA table copy (nvdisasm attributes the back edge to the statement that follows the wait loop)"
        (allowed if why else findings).append(dict(c, why=why))
    return len(ins), findings, allowed


def main():
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    lines = ["# Static constant-time audit of the secret-scalar kernels (SASS of the shipped build)", "",
             "Generated by `tools/ct_sass_audit.py` from `rustcrypto-elliptic-curves_b200/_build/curve_*.o` (nvdisasm -g).  Every conditional control transfer of",
             "`k_mul_var<C, true>`, `k_mul_gen_smem<C, true>` and `k_sign_finish<C>` (and the device functions they call) with its source line;",
             "a branch is accepted only on a line that tests public quantities (see ALLOW in the tool).  Reference discipline:",
             "`k256/src/arithmetic/mul.rs:92-127`, `primeorder/src/projective.rs:127-147`.", ""]
    total_findings = 0
    with tempfile.TemporaryDirectory() as tmp:
        for curve in CURVES:
            text = disasm(curve, tmp)
            if text is None:
                lines.append("* %s: object not built" % curve)
                continue
            sections = re.split(r"^//-+ \.text\.(\S+) -+$", text, flags=re.M)
            for k in range(1, len(sections), 2):
                name, body = sections[k], sections[k + 1].splitlines()
                if not any(re.search(p, name) for p in KERNELS):
                    continue
                dem = sh(["cu++filt", name]).stdout.strip() or name
                n, findings, allowed = audit_kernel(name, body)
                total_findings += len(findings)
                ops = " ".join(l for l in body)
                lines += ["## `%s`" % dem.replace("ecb::", ""), "",
                          "%d instructions, %d conditional control transfers / predicated memory accesses: %d on public-quantity lines, **%d findings**; indirect branches (BRX/JMX): %d" % (
                              n, len(findings) + len(allowed), len(allowed), len(findings), len(re.findall(r"\b(BRX|JMX)\b", ops))), ""]
                if allowed:
                    lines += ["| address | instruction | source line | why public | predicate set by |", "|---|---|---|---|---|"]
                    for c in allowed:
                        lines.append("| %04x | `%s` | %s:%d `%s` | %s | `%s` |" % (c["addr"], c["text"], os.path.basename(c["file"]), c["line"], c["src"][:70].replace("|", "\\|"), c["why"], c["setter"].replace("|", "\\|")))
                    lines.append("")
                for c in findings:
                    lines.append("* **FINDING** %04x `%s` at %s:%d `%s` (predicate set by `%s`) %s" % (c["addr"], c["text"], os.path.basename(c["file"]), c["line"], c["src"][:100], c["setter"], c["chain"]))
                if findings:
                    lines.append("")
    lines += ["", "**Total findings: %d**" % total_findings, ""]
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text)
    print(text if not out else "\n".join(l for l in lines if l.startswith("## ") or "FINDING" in l or "Total" in l or "conditional control" in l))
    return 1 if total_findings else 0


if __name__ == "__main__":
    sys.exit(main())
