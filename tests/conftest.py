import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    return {n: load_golden(n) for n in ("group", "field", "ecdsa", "wycheproof", "misc", "next")}


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """A fresh checkout has no libecb200.so (built artefacts are git-ignored): build it once per session when nvcc is here
    (cross-compiles for sm_100a without a GPU), exactly as __graft_entry__.build() does.  On the GPU box the library built in
    the container travels with the snapshot, so nothing happens there."""
    import importlib
    import shutil
    pkg_dir = os.path.join(ROOT, "rustcrypto-elliptic-curves_b200")
    if not os.path.exists(os.path.join(pkg_dir, "libecb200.so")) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        importlib.import_module("rustcrypto-elliptic-curves_b200.build").build(force=False, verbose=False)
    yield
