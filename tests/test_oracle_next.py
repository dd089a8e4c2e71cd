"""CPU: the oracle restatements for the SURVEY 8(f) rows (oracle/next_oracle.c) -- Stateful
derivatives of the springy Point and the rigid-body State, and the SPH neighbour pass -- against
hand-derived known answers and an independent numpy float32 restatement."""
import ctypes as C

import numpy as np

f32 = np.float32


def _deriv(orc, name, s):
    s = np.ascontiguousarray(s, f32).reshape(-1)
    d = np.empty_like(s)
    getattr(orc.lib, "orc_deriv_" + name)(s.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                                          C.c_size_t(s.size), None)
    return d


def test_springy_point_derivative_known_answer(orc):
    # springy_mesh.rs:223-240: mass 2, velocity (1,0,0), accumulated force (2,-4,6)
    d = _deriv(orc, "springy_point", [2, 5, 6, 7, 1, 0, 0, 2, -4, 6])
    assert d.tolist() == [0, 1, 0, 0, 1, -2, 3, 0, 0, 0]
    out = orc.state_euler(np.array([2, 5, 6, 7, 1, 0, 0, 2, -4, 6], f32), 0.5, "springy_point")
    assert out.tolist() == [2, 5.5, 6, 7, 1.5, -1, 1.5, 2, -4, 6]


def _rigid_numpy(s):
    """rigidbody.rs:103-140 in numpy float32, written from the Rust source (cgmath order)."""
    s = s.astype(f32)
    q = s[3:6]; qs = s[6]; m = s[13]
    d = np.zeros(29, f32)
    d[0:3] = s[7:10] / m
    x2, y2, z2 = q + q
    xx2, xy2, xz2 = x2 * q[0], x2 * q[1], x2 * q[2]
    yy2, yz2, zz2 = y2 * q[1], y2 * q[2], z2 * q[2]
    sy2, sz2, sx2 = y2 * qs, z2 * qs, x2 * qs
    one = f32(1)
    R = np.array([[one - yy2 - zz2, xy2 + sz2, xz2 - sy2],        # columns
                  [xy2 - sz2, one - xx2 - zz2, yz2 + sx2],
                  [xz2 + sy2, yz2 - sx2, one - xx2 - yy2]], f32)
    I0 = s[14:23].reshape(3, 3)                                    # columns

    def mul(a, b):                                                 # a, b: [col][row]
        o = np.zeros((3, 3), f32)
        for c in range(3):
            for r in range(3):
                o[c, r] = f32(f32(a[0, r] * b[c, 0]) + f32(a[1, r] * b[c, 1])) + f32(a[2, r] * b[c, 2])
        return o
    Iinv = mul(mul(R, I0), R.T.copy())
    L = s[10:13]
    w = np.array([f32(f32(Iinv[0, r] * L[0]) + f32(Iinv[1, r] * L[1])) + f32(Iinv[2, r] * L[2]) for r in range(3)], f32)
    a = f32(0.5) * w
    az = f32(0.5) * f32(0.0)
    d[6] = az * qs - a[0] * q[0] - a[1] * q[1] - a[2] * q[2]
    d[3] = az * q[0] + a[0] * qs + a[1] * q[2] - a[2] * q[1]
    d[4] = az * q[1] + a[1] * qs + a[2] * q[0] - a[0] * q[2]
    d[5] = az * q[2] + a[2] * qs + a[0] * q[1] - a[1] * q[0]
    d[7:10] = s[23:26]
    d[10:13] = s[26:29]
    return d


def test_rigid_body_derivative(orc):
    one = np.zeros(29, f32)
    one[6] = 1; one[13] = 2; one[[14, 18, 22]] = 1
    one[7:10] = [2, 4, 6]; one[10:13] = [0.2, 0.4, 0.6]; one[23:29] = [1, 2, 3, 4, 5, 6]
    d = _deriv(orc, "rigidbody", one)
    assert d[:3].tolist() == [1, 2, 3]
    assert np.array_equal(d[3:7], np.array([0.1, 0.2, 0.3, 0.0], f32))   # 0.5 (0, L) * identity
    assert d[7:13].tolist() == [1, 2, 3, 4, 5, 6] and not d[13:].any()
    rng = np.random.default_rng(8)
    for _ in range(200):
        s = rng.normal(size=29).astype(f32)
        s[3:7] /= np.linalg.norm(s[3:7])
        s[13] = abs(s[13]) + f32(0.5)
        assert np.array_equal(_deriv(orc, "rigidbody", s).view(np.uint32), _rigid_numpy(s).view(np.uint32))


def test_sph_neighbours_lattice_and_brute_force(orc):
    g = np.arange(-4, 4, dtype=np.float32) * f32(0.1)
    lattice = np.array([[x, y, z] for x in g for z in g for y in g], f32)
    idx, cnt, den = orc.sph_neighbors(lattice, 8, 0.25, 0.001)
    # every particle is its own nearest neighbour (d2 = 0) and finds 8 within 0.25 (2.5 spacings)
    assert np.array_equal(idx[:, 0], np.arange(512, dtype=np.uint32)) and (cnt == 8).all()
    rng = np.random.default_rng(9)
    pos = (rng.random((400, 3)) * 0.8).astype(f32)
    idx, cnt, den = orc.sph_neighbors(pos, 8, 0.1, 0.001)
    d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).astype(f32)
    d2 = ((f32(0) + d2[..., 0]) + d2[..., 1]) + d2[..., 2]
    for i in range(400):
        order = np.lexsort((np.arange(400), d2[i]))
        want = [j for j in order[:8] if d2[i, j] < f32(0.1) * f32(0.1)]
        assert cnt[i] == len(want) and idx[i, :cnt[i]].tolist() == want
        assert (idx[i, cnt[i]:] == 0xFFFFFFFF).all()
    # monaghan(0, s) = 1 / (pi s^3): a lone particle's density is mass times that
    _, c1, d1 = orc.sph_neighbors(np.zeros((1, 3), f32), 8, 0.1, 0.001)
    assert c1[0] == 1 and d1[0] == f32(0.001) * (f32(1.0) / (f32(np.pi) * (f32(0.1) * f32(0.1) * f32(0.1))))
