"""Build libecb200 with extra -D flags into rustcrypto-elliptic-curves_b200/variants/libecb200_<name>.so (A/B experiments on
the GPU box: ECB200_LIB=<path> python scripts/quick_bench.py ...).  usage: build_variant.py name -DFOO=1 [-DBAR=2 ...]"""
import concurrent.futures, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rustcrypto-elliptic-curves_b200")
sys.path.insert(0, PKG)
import build as b  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
obj = os.path.join(PKG, "variants", "_obj_" + name)
os.makedirs(obj, exist_ok=True)


def cc(unit):
    o = os.path.join(obj, unit + ".o")
    r = subprocess.run(["nvcc"] + b.NVCC_FLAGS + defs + ["-c", os.path.join(b.CSRC, unit + ".cu"), "-o", o], capture_output=True, text=True)
    open(os.path.join(obj, unit + ".ptxas.log"), "w").write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError(r.stderr[-3000:])
    return o


with concurrent.futures.ThreadPoolExecutor(max_workers=5) as ex:
    objs = list(ex.map(cc, b.UNITS))
lib = os.path.join(PKG, "variants", "libecb200_%s.so" % name)
subprocess.check_call(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
print("built", lib)
