"""Host-side logic of the API mirror (no GPU): Duration, Parametric/LeadBoid
stepping (boid.rs:46-53, parametric.rs:17-21) against the oracle, synthetic
input generator."""
import numpy as np
import pytest

from feriphys_b200 import synth
from feriphys_b200.flocking import (DEMO_PATHS, Config, Duration, LeadBoid, Panic, Parametric)

f32 = np.float32


def test_duration_matches_oracle(orc):
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(0, 10, 500), 10.0 ** rng.uniform(-11, 10, 500)]).astype(f32)
    for x in xs:
        f, s, n = orc.duration_from_secs_f32(float(x))
        d = Duration.from_secs_f32(x)
        assert f == 0 and (d.secs, d.nanos) == (s, n)
        assert d.as_secs_f32() == f32(orc.duration_as_secs_f32(s, n))
    for bad in (-1e-3, float("nan"), float("inf"), 2.0 ** 64):
        with pytest.raises(Panic):
            Duration.from_secs_f32(bad)
    assert Duration.from_secs_f32(-0.0) == Duration(0, 0)
    assert Duration.from_millis(1).as_secs_f32() == f32(0.001)
    assert Duration.from_secs(4) > Duration(3, 999_999_999)


def test_parametric_returns_then_advances():
    p = Parametric(lambda t: (t, f32(2) * t, f32(1)))
    assert list(p.step(f32(0.5))) == [0.0, 0.0, 1.0]
    assert p.curr_time == f32(0.5)
    assert list(p.step(f32(0.5))) == [0.5, 1.0, 1.0]


def test_lead_boid_rows_match_oracle_for_demo_paths(orc):
    dt = Config().dt
    for kinds in ([0], [1, 2]):
        leads = [LeadBoid(DEMO_PATHS[k]) for k in kinds]
        o_leads, o_times = orc.make_leads(kinds)
        for step in range(300):
            rows = np.stack([l.row() for l in leads])
            assert np.array_equal(rows.view(np.uint32), o_leads.view(np.uint32)), (kinds, step)
            for l in leads:
                l.step(Duration.from_secs_f32(dt))
            orc.step_leads(o_leads, o_times, kinds, float(f32(dt)))
        assert leads[0].parametric.curr_time == o_times[0]


def test_synth_is_deterministic_and_shaped():
    a = synth.uniform_flock(1000, 24.0, seed=9)
    b = synth.uniform_flock(1000, 24.0, seed=9)
    assert np.array_equal(a, b) and a.dtype == np.float32 and a.shape == (1000, 6)
    assert a[:, :3].min() >= 0 and a[:, :3].max() < 24 and a[:, 3:].min() >= -1 and a[:, 3:].max() < 1
    # index-keyed: a rank can generate its own slice of the global flock
    c = synth.uniform_flock(400, 24.0, seed=9, first=600)
    assert np.array_equal(c, a[600:])
    assert not np.array_equal(a, synth.uniform_flock(1000, 24.0, seed=10))
    s = synth.spawn_flock([(15, 10, 0), (25, 0.5, 0)], 111)
    assert s.shape == (110, 6)            # integer division drops the remainder (flocking.rs:76)
    assert np.all(s[:55, 0] >= 15) and np.all(s[:55, 0] < 16) and np.all(s[55:, 0] >= 25)
    u = synth.u01(1, 100000, 1)
    assert abs(float(u.mean()) - 0.5) < 0.01
