"""Builds libecb200.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

    python -m importlib ... or: python rustcrypto-elliptic-curves_b200/build.py [--force]

One translation unit per curve + the host runtime, compiled in parallel, linked with -shared.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libecb200.so")
UNITS = ["abi", "curve_k256", "curve_p256", "curve_p384", "curve_sm2", "curve_p192", "curve_p224"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "ecb200.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


MOV_PATCH = os.environ.get("ECB200_MOV_PATCH", "0") != "0"      # tools/sass_mov_patch.py between ptxas and fatbinary (curve units)
PATCHER = os.path.join(os.path.dirname(HERE), "tools", "sass_mov_patch.py")


def _compile_steps(unit, src, obj, objdir, flags, mov_patch):
    """nvcc's own sub-commands (`nvcc -dryrun`: preprocess, cudafe++, cicc, ptxas, fatbinary, host compile) run one by one, with
    the SASS pass of tools/sass_mov_patch.py applied to the cubin between ptxas and fatbinary.  Returns the combined log."""
    import re
    import shlex
    tmp = os.path.join(objdir, "_tmp_" + unit)
    os.makedirs(tmp, exist_ok=True)
    dry = subprocess.run(["nvcc", "-dryrun", "--keep", "--keep-dir", tmp] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
    if dry.returncode != 0:
        raise RuntimeError("nvcc -dryrun failed for %s:\n%s" % (unit, dry.stderr[-2000:]))
    env = dict(os.environ)
    log = []
    for line in dry.stderr.splitlines():
        if not line.startswith("#$ "):
            continue
        cmd = line[3:].strip()
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)=(.*)$", cmd)
        if m and " " not in m.group(1) and not cmd.startswith(("gcc", "ptxas", "cicc", "cudafe", "fatbinary", "rm ", "\"")):
            val = m.group(2).strip()
            env[m.group(1)] = os.path.expandvars(val.strip('"')) if m.group(1) not in ("INCLUDES", "LIBRARIES") else val
            continue
        r = subprocess.run(cmd, shell=True, capture_output=True, text=True, env=env)
        log.append("$ " + cmd[:300] + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("build step failed for %s:\n%s\n%s" % (unit, cmd[:500], (r.stdout + r.stderr)[-4000:]))
        if cmd.startswith("ptxas ") and mov_patch:
            cubin = shlex.split(cmd)[shlex.split(cmd).index("-o") + 1]
            pr = subprocess.run([sys.executable, PATCHER, cubin, cubin + ".mov"], capture_output=True, text=True)
            log.append("$ sass_mov_patch %s\n%s%s" % (os.path.basename(cubin), pr.stdout, pr.stderr))
            if pr.returncode != 0:
                raise RuntimeError("sass_mov_patch failed for %s:\n%s" % (unit, (pr.stdout + pr.stderr)[-3000:]))
            os.replace(cubin + ".mov", cubin)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return "\n".join(log)


def _compile(unit, objdir=None, extra=(), mov_patch=None):
    objdir = objdir or OBJ
    mov_patch = MOV_PATCH if mov_patch is None else mov_patch
    src = os.path.join(CSRC, unit + ".cu")
    obj = os.path.join(objdir, unit + ".o")
    log = os.path.join(objdir, unit + ".ptxas.log")
    flags = NVCC_FLAGS + list(extra)
    if mov_patch and unit.startswith("curve_"):
        text = _compile_steps(unit, src, obj, objdir, flags, True)
        with open(log, "w") as f:
            f.write("nvcc (stepwise, with tools/sass_mov_patch.py) " + " ".join(flags) + "\n" + text)
        return obj
    cmd = ["nvcc"] + flags + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (unit, (r.stdout + r.stderr)[-4000:]))
    return obj


def build_variant(name, defs=(), mov_patch=False):
    """libecb200_<name>.so under variants/ with extra -D flags and / or the SASS pass switched (A/B runs: ECB200_LIB=<path>)."""
    objdir = os.path.join(HERE, "variants", "_obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(lambda u: _compile(u, objdir, defs, mov_patch), UNITS))
    lib = os.path.join(HERE, "variants", "libecb200_%s.so" % name)
    subprocess.check_call(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return lib


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(_compile, UNITS))
    cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
