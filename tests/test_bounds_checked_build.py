"""compute-sanitizer is not available on the GPU pool, so the kernels carry their own range checks in a
-DCLB_BOUNDS_CHECK build (variants/lib_bounds.so, built by __graft_entry__.build()): parity workloads run under it and
the device-side violation counter must stay at zero."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "variants", "lib_bounds.so")


@pytest.mark.gpu
def test_no_out_of_bounds_access_in_the_checked_build():
    if not os.path.exists(LIB):
        pytest.skip("variants/lib_bounds.so has not been built (python -c 'import __graft_entry__ as g; g.build()')")
    env = dict(os.environ, CLB_LIB=LIB)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bounds_check_run.py")], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1200)
    assert p.returncode == 0, p.stderr[-2000:]
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert out["bounds_checked_build"] is True and out["out_of_bounds_accesses_caught"] == 0 and out["parity_cases"] >= 20
