#!/usr/bin/env python
"""Run the REFERENCE'S OWN test-suite (/root/reference/tests: test_standard, test_square_root, test_rodeofor,
test_rodeojit, test_fitz, test_add_sqrt) with oracle/jaxshim standing in for jax.

This validates the stand-in itself: the brute-force-conditioning known-answer tests of the Kalman primitives, the
scan == for-loop tests of the solvers, the jit == eager tests and the FitzHugh-Nagumo-vs-odeint test all pass on it
(28 passed, 2026-10-18), so the vectors that tests/golden/make_reference_golden.py produces through the same stand-in
come from an execution of the reference that the reference's own tests accept.  Build-container only (needs
/root/reference); nothing under tests/ imports this file.

    python tests/golden/run_reference_tests_over_shim.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("RODEO_REFERENCE", "/root/reference")

if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        tests = os.path.join(tmp, "reftests")
        shutil.copytree(os.path.join(REF, "tests"), tests)          # /root/reference is read-only: pytest needs a writable cwd
        env = dict(os.environ, PYTHONPATH=os.pathsep.join(
            [os.path.join(ROOT, "oracle", "jaxshim"), os.path.join(REF, "src"), tests]))
        sys.exit(subprocess.call([sys.executable, "-W", "ignore", "-m", "pytest", "-q", "-p", "no:cacheprovider", "."],
                                 cwd=tests, env=env))
