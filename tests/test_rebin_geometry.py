"""CPU property test of the grid path's exactness argument, with the same
float32 arithmetic as the kernels (fp_grid.cuh: cell_coord = floor(fl(fl(x - origin) * inv)),
clamped): for a cell edge of reach * (1 + 1/512) + skin, sliced zspan times along z, and boids
that have drifted by up to skin / 2 from where they were binned, EVERY pair closer than `reach`
has home cells at most one apart in x and y and at most zspan slices apart in z -- so the walk
over the 27 cells (9 rows of 2 * zspan + 1 slices) around the home cell sees every neighbour,
including boids clamped into edge cells from outside the fitted domain."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

f32 = np.float32


def _coord(x, origin, inv, dim):
    c = np.floor((x.astype(f32) - f32(origin)).astype(f32) * f32(inv)).astype(np.int64)
    return np.clip(c, 0, dim - 1)


@pytest.mark.parametrize("reach,skin,zspan,extent,offset", [
    (16.0, 0.0, 1, 300.0, 0.0),
    (16.0, 0.11, 4, 300.0, 0.0),
    (16.0, 2.0, 4, 500.0, 1000.0),       # large coordinates: coarser float32 spacing
    (5.0, 0.6, 2, 90.0, -45.0),
    (30.0, 0.25, 4, 2048.0, 0.0),
])
def test_in_range_pairs_stay_within_the_home_neighbourhood(reach, skin, zspan, extent, offset):
    rng = np.random.default_rng(int(reach * 1000 + skin * 100 + zspan))
    n = 60000
    binned = (rng.random((n, 3)) * extent + offset).astype(f32)
    # the grid is fitted to the bulk only: a tenth of the boids sit outside and are clamped
    lo = binned.min(axis=0) + f32(0.05 * extent)
    hi = binned.max(axis=0) - f32(0.05 * extent)
    cell = reach * (1.0 + 1.0 / 512.0) + skin
    dims = [int(np.floor((float(hi[a]) - float(lo[a])) / cell)) + 1 for a in range(3)]
    dimz = int(np.floor((float(hi[2]) - float(lo[2])) / (cell / zspan))) + 1
    inv = f32(1.0) / f32(cell)
    invz = f32(zspan / cell)
    cx = _coord(binned[:, 0], lo[0], inv, dims[0])
    cy = _coord(binned[:, 1], lo[1], inv, dims[1])
    cz = _coord(binned[:, 2], lo[2], invz, dimz)
    # drift: up to skin / 2 in a random direction (the device bound is on the Euclidean norm)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= (rng.random((n, 1)) ** (1 / 3)) * (skin / 2)
    cur = (binned.astype(np.float64) + d).astype(f32)
    # every pair within reach NOW (float64 distance, a hair generous)
    pairs = cKDTree(cur.astype(np.float64)).query_pairs(reach * (1 + 1e-6), output_type="ndarray")
    assert len(pairs) > 10000
    i, j = pairs[:, 0], pairs[:, 1]
    assert np.abs(cx[i] - cx[j]).max() <= 1
    assert np.abs(cy[i] - cy[j]).max() <= 1
    assert np.abs(cz[i] - cz[j]).max() <= zspan
    # and the bound is not vacuous: with eight times the allowed drift some neighbour is missed
    if skin >= 0.5:
        far = (binned.astype(np.float64) + 8 * d).astype(f32)
        pf = cKDTree(far.astype(np.float64)).query_pairs(reach * (1 - 1e-6), output_type="ndarray")
        a, b = pf[:, 0], pf[:, 1]
        worst = max(np.abs(cx[a] - cx[b]).max(), np.abs(cy[a] - cy[b]).max(),
                    (np.abs(cz[a] - cz[b]).max() + zspan - 1) // zspan)
        assert worst >= 2


@pytest.mark.parametrize("reach,skin,extent,offset", [
    (16.0, 0.11, 300.0, 0.0),
    (16.0, 2.0, 500.0, 1000.0),
    (5.0, 0.6, 90.0, -45.0),
])
def test_candidate_list_cut_keeps_every_pair_that_can_come_into_reach(reach, skin, extent, offset):
    """The standing candidate lists (fp_walk_nl.cu) keep, at binning time, the pairs
    whose FUSED float32 squared distance is below m2_wide = (reach + skin)^2 (1 + 1e-5).  While
    every boid stays within skin / 2 of its binned position, every pair the step's pre-gate can
    keep -- fused float32 squared distance below m2_cut (1 + 1e-6) -- must be among them."""
    rng = np.random.default_rng(int(reach * 1000 + skin * 100))
    n = 60000
    binned = (rng.random((n, 3)) * extent + offset).astype(f32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= (skin / 2) * np.where(rng.random((n, 1)) < 0.5, 1.0, rng.random((n, 1)))  # half of them at the limit
    cur = (binned.astype(np.float64) + d).astype(f32)
    R = float(f32(reach)) + float(f32(skin))
    m2_wide = np.nextafter(f32(R * R * (1.0 + 1e-5)), f32(np.inf))
    m2_cut_hi = f32(float(f32(reach)) ** 2 * (1.0 + 2e-6))    # at least the kernels' m2_cut (1 + 1e-6)

    def fused_m2(p, i, j):                                   # fma(dz, dz, fma(dy, dy, dx * dx)) in float32
        dd = (p[j] - p[i]).astype(np.float64)                # the float32 differences, exactly
        m = f32(dd[:, 0] * dd[:, 0])
        m = (dd[:, 1] * dd[:, 1] + m.astype(np.float64)).astype(f32)
        return (dd[:, 2] * dd[:, 2] + m.astype(np.float64)).astype(f32)

    pairs = cKDTree(cur.astype(np.float64)).query_pairs(reach * (1 + 1e-4), output_type="ndarray")
    i, j = pairs[:, 0], pairs[:, 1]
    now = fused_m2(cur, i, j)
    kept_now = now < m2_cut_hi
    assert kept_now.sum() > 10000
    then = fused_m2(binned, i[kept_now], j[kept_now])
    assert (then < m2_wide).all()
    # not vacuous: the closest call is within a few percent of the cut
    assert float(then.max()) > 0.9 * float(m2_wide)


@pytest.mark.parametrize("reach,skin,extent", [(16.0, 0.11, 300.0), (16.0, 2.0, 500.0), (5.0, 0.6, 90.0)])
def test_lists_built_mid_binning_need_twice_the_skin(reach, skin, extent):
    """Variant 46 builds the lists at a binning's second step, from the positions of that moment:
    both then and later every boid is within skin / 2 of its BINNED position, so a pair within
    reach later is within reach + 2 skin at the build -- and reach + skin is not enough."""
    rng = np.random.default_rng(int(reach * 10 + skin * 1000))
    n = 60000
    binned = rng.random((n, 3)) * extent

    def drift():
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        return d * (skin / 2) * np.where(rng.random((n, 1)) < 0.5, 1.0, rng.random((n, 1)))

    at_build = (binned + drift()).astype(f32)
    later = (binned + drift()).astype(f32)
    pairs = cKDTree(later.astype(np.float64)).query_pairs(reach * (1 + 1e-6), output_type="ndarray")
    i, j = pairs[:, 0], pairs[:, 1]
    d = (at_build[j] - at_build[i]).astype(np.float64)
    m2 = (d * d).sum(axis=1)
    wide = lambda skins: (reach + skins * skin) ** 2 * (1.0 + 1e-5)
    assert (m2 * (1 + 1e-6) < wide(2.0)).all()
    if skin >= 0.5:
        assert (m2 >= wide(1.0)).any()
