"""-m gpu, OPT-IN (set FP_TEST_EXPERIMENTAL=1): the walk on standing candidate lists.

FP_WALK_VARIANT=41 selects fp_walk_nl.cu (candidate lists built once per binning) instead of
the production walk for the steps of a single-GPU grid flock.  It is not the default and was
written when no GPU time was left to try it, so these checks do not run unless asked for.  The
variant is read from the environment when the library is first used, hence the subprocesses.

What must hold: every -m gpu grid test passes unchanged (they compare with the oracle bit for
bit while a binning stands), and whole runs agree bit for bit with the production kernel --
including a flock dense enough to overflow the lists, which must fall back by itself."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FP_TEST_EXPERIMENTAL") != "1",
                                 reason="experimental kernel variant: set FP_TEST_EXPERIMENTAL=1")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _env(variant):
    return dict(os.environ, FP_WALK_VARIANT=str(variant), FP_NL_TRACE="1")


def test_grid_suite_passes_on_candidate_lists():
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-s", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_gpu_rebin.py"), os.path.join(ROOT, "tests", "test_gpu_grid.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800, env=_env(41), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "candidate lists on" in r.stderr, "the variant was never used"


def _hash(variant, *args):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "nl_state_hash.py"), *map(str, args)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=_env(variant), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1]), r.stderr


@pytest.mark.parametrize("args", [(200_000, 470.0, 120, 7), (1 << 20, 816.0, 60, 11)])
def test_runs_agree_bit_for_bit_with_the_production_walk(args):
    a, _ = _hash(31, *args)
    b, err = _hash(41, *args)
    assert "candidate lists on" in err and "overflowed" not in err
    assert a["finite"] and a["rebins"] == b["rebins"] and a["replayed"] == b["replayed"]
    assert a["sha256"] == b["sha256"]


def test_overflowing_lists_fall_back_to_the_production_walk():
    args = (60_000, 315.0, 40, 5, 6000)      # a 6000-boid ball of radius 6: thousands of neighbours each
    a, _ = _hash(31, *args)
    b, err = _hash(41, *args)
    assert "overflowed" in err
    assert a["finite"] and a["sha256"] == b["sha256"]
