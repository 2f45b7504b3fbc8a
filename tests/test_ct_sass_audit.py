"""CPU tier: the static constant-time audit of the secret-scalar kernels (tools/ct_sass_audit.py) must report no conditional
branch outside the public-quantity allow list, on the SASS of the library that ships (north_star: "no secret-dependent
branches"; k256/src/arithmetic/mul.rs:92-127, primeorder/src/projective.rs:127-147).  Needs nvdisasm / cuobjdump (CUDA toolkit)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not (shutil.which("nvdisasm") and shutil.which("cuobjdump")), reason="CUDA binary utilities not installed")
def test_secret_scalar_kernels_have_no_data_dependent_branch(tmp_path):
    import __graft_entry__
    __graft_entry__.build()          # objects under rustcrypto-elliptic-curves_b200/_build (no-op when up to date)
    out = tmp_path / "audit.md"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ct_sass_audit.py"), "--out", str(out)], capture_output=True, text=True)
    text = out.read_text()
    assert r.returncode == 0, r.stdout[-3000:]
    assert "Total findings: 0" in text
    # all three kernel families of all six curves were actually inspected
    assert text.count("## `void k_mul_var<") == 6 and text.count("## `void k_mul_gen_smem<") == 6 and text.count("## `void k_sign_finish<") == 6
    assert text.count("## `void k_gen_half<") == 6 and text.count("## `void k_sum_normalize<") == 6      # split fixed-base path
    assert "indirect branches (BRX/JMX): 0" in text and "BRX/JMX): 1" not in text
