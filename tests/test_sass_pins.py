"""The kernels whose GPU results are quoted (profiles/) are byte-identical to the
build that was last run on a B200.  Most edits happen without a GPU; this is what lets a
refactor claim it left a measured kernel alone.  After changing a kernel on purpose AND
re-running it on hardware, pin the new build: python tools/sass_pins.py --record"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_measured_kernels_are_unchanged():
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    if not os.path.isdir(os.path.join(ROOT, "feriphys_b200", "csrc", "_build")):
        pytest.skip("library not built in-tree")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_pins.py")], capture_output=True, text=True)
    if r.returncode == 2:
        pytest.skip(r.stdout.strip())
    assert r.returncode == 0, ("kernels differ from the build last run on hardware -- re-validate on a B200, then "
                               "`python tools/sass_pins.py --record`:\n" + r.stdout[-3000:] + r.stderr[-1000:])
