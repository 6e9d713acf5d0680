"""The oracle against the UNMODIFIED reference binary on randomized scenes (oracle/fuzz_oracle_vs_ref.py): pinning beyond the
committed golden vectors. Needs oracle/_ref/ref_mpm, which `make -C oracle ref` builds from /root/reference where that
exists (this container) and which travels with the repo snapshot; skipped where it is absent."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_mpm")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_mpm not built (no /root/reference here)")
def test_oracle_matches_the_unmodified_reference_on_random_scenes():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "fuzz_oracle_vs_ref.py"), "16", "100"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200, cwd=ROOT)
    assert r.returncode == 0 and "ORACLE_VS_REFERENCE_OK 16 cases" in r.stdout, r.stdout[-4000:]
