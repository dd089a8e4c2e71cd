"""-m gpu: the grid step on standing candidate lists (fp_walk_nl.cu, the default) against the plain
staged walk (FP_NL=0, fp_walk.cu).

Under EXACT numerics the two must agree bit for bit whenever they keep the same binnings (the skin
is pinned with FP_SKIN for that): whole runs, a flock dense enough to overflow the lists -- which
must fall back by itself -- and every stage of tools/nl_transitions.py (taps, replays, config and
state changes).  The switch is read from the environment when the library is first used, hence the
subprocesses."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _env(lists, **extra):
    env = dict(os.environ, FP_NL="1" if lists else "0", FP_NL_TRACE="1", FP_SKIN="0.2", FP_NUMERICS="exact")
    env.update(extra)
    return env


def _hash(lists, *args):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "nl_state_hash.py"), *map(str, args)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=_env(lists), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1]), r.stderr


@pytest.mark.parametrize("args", [(200_000, 470.0, 120, 7), (1 << 20, 816.0, 400, 11)])
def test_runs_agree_bit_for_bit_with_the_staged_walk(args):
    a, err0 = _hash(False, *args)
    b, err = _hash(True, *args)
    assert "candidate lists on" in err and "candidate lists off" not in err
    assert "candidate lists on" not in err0
    assert a["finite"] and a["rebins"] == b["rebins"] and a["replayed"] == b["replayed"]
    assert a["rebins"] >= 3
    assert a["sha256"] == b["sha256"]


def test_overflowing_lists_fall_back_to_the_staged_walk():
    # a 6000-boid ball of radius 6, thousands of neighbours each: its CTAs get no lists (they walk
    # from global memory), and at the next binning the library drops the lists altogether
    args = (60_000, 315.0, 80, 5, 6000)
    a, _ = _hash(False, *args)
    b, err = _hash(True, *args)
    assert "CTAs without candidate lists" in err
    assert a["finite"] and a["rebins"] == b["rebins"] and a["sha256"] == b["sha256"]


def test_state_transitions_agree_with_the_staged_walk():
    outs = []
    for lists in (False, True):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "nl_transitions.py")],
                           capture_output=True, text=True, timeout=900, env=_env(lists), cwd=ROOT)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append([ln for ln in r.stdout.splitlines() if ln[:2].strip().isdigit()])
    assert len(outs[0]) == 10 and outs[0] == outs[1]
