"""Memory-safety substitute for compute-sanitizer (closed on this pool): tools/memsafety_run.py drives the builder, the
pooling kernels in every regime, the heavy-cell kernels and the fused conv through libshpl_debug.so (in-kernel index
checks, `make debug`) on canary-padded buffers.  Run in a subprocess because this process has the product library loaded."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_debug_build_with_canaries_reports_no_out_of_bounds_access():
    dbg = os.path.join(ROOT, "sparse_pooling_b200", "libshpl_debug.so")
    if not os.path.exists(dbg):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "sparse_pooling_b200", "csrc"), "debug"])
    env = dict(os.environ, SHPL_LIB=dbg)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "memsafety_run.py")], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=900)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0 and "memsafety: PASS" in r.stdout
