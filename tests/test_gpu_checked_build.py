"""The pool has no compute-sanitizer (profiles/r02b_compute_sanitizer_closed.log), so the memory-safety check of the
kernels is a build of the same library with device-side bounds assertions (`make checked`, -DHMG_BOUNDS: every read of
the shared-memory ring, every output store of the apply kernel, every entry the interface and CG kernels address): the
parity tests of the deepest small cases run through it in a child process.  A failed assertion aborts the child."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "variants", "libhmg_checked.so")


def test_parity_suite_through_the_bounds_checked_library():
    if not os.path.exists(CHECKED):
        pytest.skip("variants/libhmg_checked.so not built (make checked)")
    env = dict(os.environ, HMG_LIB=CHECKED)
    sel = "(tet-c2-L5 or tri-c3-L7 or tet-c2-L6 or tet-c3-L3-ordered) and (global_product or smoothing or interface or residual_history or mul_matches)"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q", "-k", sel],
                       env=env, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
