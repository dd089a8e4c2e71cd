"""The oracle against the reference's own golden tests for the State module:
state.rs:166-185 (euler_step, exact equality) and state.rs:218-280 (rk4_step,
+-0.005 against the listed values, exact t and carried constant)."""
import numpy as np

f32 = np.float32


def test_euler_step_state_rs_166_185(orc):
    s = np.array([0, 0, 0, 0, 0, 1], f32)
    out = orc.state_euler(s, 0.5, "point")
    assert list(out[:3]) == [0.0, 0.0, 0.5]
    assert list(out[3:]) == [0.5, -0.5, 1.0]


def test_rk4_step_state_rs_218_280(orc):
    s = np.array([0.5, 0.0, 0.5], f32)
    golden = [1.425130208333333, 2.640859085770477, 4.009155464830968, 5.305471950534675]
    ts = [0.5, 1.0, 1.5, 2.0]
    ys = []
    for g, t in zip(golden, ts):
        s = orc.state_rk4(s, 0.5, "examplefn")
        assert g - 0.005 < s[0] < g + 0.005
        assert s[1] == t and s[2] == 0.5
        ys.append(float(s[0]))
    # values the survey re-computed in numpy float32
    assert np.allclose(ys, [1.4251302, 2.6396027, 4.006819, 5.301605], rtol=0, atol=2e-6)


def test_python_callback_derivative_matches_builtin(orc):
    def deriv(s):
        return np.array([s[0] - s[1] * s[1] + f32(1.0), 1.0, 0.0], f32)
    s = np.array([0.5, 0.0, 0.5], f32)
    assert np.array_equal(orc.state_rk4(s, 0.5, deriv), orc.state_rk4(s, 0.5, "examplefn"))
    assert np.array_equal(orc.state_euler(s, 0.5, deriv), orc.state_euler(s, 0.5, "examplefn"))


def test_state_euler_is_flocking_inline_euler_F1(orc):
    # State::euler_step (x*h then s+delta) rounds like flocking.rs:116-117
    rng = np.random.default_rng(0)
    st = rng.normal(size=(50, 6)).astype(f32)
    acc = rng.normal(size=(50, 3)).astype(f32)
    h = f32(0.001)

    def deriv(s):
        s = s.reshape(-1, 6)
        return np.concatenate([s[:, 3:], acc], axis=1).reshape(-1)
    out = orc.state_euler(st.reshape(-1), float(h), deriv).reshape(-1, 6)
    assert np.array_equal(out[:, :3], st[:, :3] + h * st[:, 3:])
    assert np.array_equal(out[:, 3:], st[:, 3:] + h * acc)
