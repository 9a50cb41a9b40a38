"""The exactness argument behind csrc/cell_uf.cuh, checked on the CPU in IEEE fp32 (numpy, no FMA): for two axis-aligned
boxes the values lower = d2(nearest gap) and upper = d2(farthest corners), computed with the reference's own operation
order ((dx*dx)+dy*dy)+dz*dz, bound the fp32 predicate value of EVERY point pair drawn from the boxes -- so
`upper < r2 => union without looking at a point` and `lower >= r2 => skip` never change the partition
(reference call site MOT.cpp:472-488)."""
import numpy as np

f32 = np.float32


def d2(ax, ay, az, bx, by, bz):
    dx, dy, dz = f32(ax - bx), f32(ay - by), f32(az - bz)
    return f32(f32(f32(dx * dx) + f32(dy * dy)) + f32(dz * dz))


def axis_bounds(alo, ahi, blo, bhi):
    d1 = (alo - bhi).astype(f32)
    d2_ = (ahi - blo).astype(f32)
    gap = np.maximum(f32(0), np.maximum(d1, -d2_))
    far = np.maximum(np.abs(d1), np.abs(d2_))
    return gap, far


def sq3(x, y, z):
    return ((x * x).astype(f32) + (y * y).astype(f32)).astype(f32) + (z * z).astype(f32)


def test_box_bounds_hold_in_fp32():
    rng = np.random.default_rng(7)
    for scale, offset in [(1.0, 0.0), (0.3, 100.0), (1e-3, 0.0), (5.0, -2000.0), (0.5, 1e5)]:
        n_box, n_pts = 400, 12
        # two sets of points per trial; boxes are their exact min / max (as k_cell_local computes them)
        a = (rng.random((n_box, n_pts, 3)) * scale + offset).astype(f32)
        shift = (rng.normal(size=(n_box, 1, 3)) * scale).astype(f32)
        b = ((rng.random((n_box, n_pts, 3)) * scale + offset).astype(f32) + shift).astype(f32)
        alo, ahi, blo, bhi = a.min(1), a.max(1), b.min(1), b.max(1)
        g, fr = zip(*[axis_bounds(alo[:, k], ahi[:, k], blo[:, k], bhi[:, k]) for k in range(3)])
        lower = sq3(*g).astype(f32)
        upper = sq3(*fr).astype(f32)
        # every pair of every trial
        pa = a[:, :, None, :]
        pb = b[:, None, :, :]
        dx = (pa[..., 0] - pb[..., 0]).astype(f32)
        dy = (pa[..., 1] - pb[..., 1]).astype(f32)
        dz = (pa[..., 2] - pb[..., 2]).astype(f32)
        val = sq3(dx, dy, dz).astype(f32)
        assert (val >= lower[:, None, None]).all(), "a pair fell below the gap bound"
        assert (val <= upper[:, None, None]).all(), "a pair exceeded the far-corner bound"
        # the bounds are attained by corner points, so they are tight as well
        assert np.array_equal(upper, np.maximum(upper, val.max((1, 2))))


def test_point_box_bound_holds_in_fp32():
    rng = np.random.default_rng(8)
    b = (rng.random((500, 9, 3)) * 0.7 + 50.0).astype(f32)
    p = (rng.random((500, 3)) * 2.0 + 49.0).astype(f32)
    lo, hi = b.min(1), b.max(1)
    g = [np.maximum(f32(0), np.maximum((lo[:, k] - p[:, k]).astype(f32), (p[:, k] - hi[:, k]).astype(f32))) for k in range(3)]
    lower = sq3(*g).astype(f32)
    dx = (p[:, None, 0] - b[..., 0]).astype(f32)
    dy = (p[:, None, 1] - b[..., 1]).astype(f32)
    dz = (p[:, None, 2] - b[..., 2]).astype(f32)
    assert (sq3(dx, dy, dz).astype(f32) >= lower[:, None]).all()


def test_fine_candidate_mask_covers_chebyshev_2():
    # axis_allowed(): children b of the neighbour at coarse offset d that can be within 2 fine cells of child a
    def allowed(d, abit):
        if d > 0:
            return {0, 1} if abit else {0}
        if d < 0:
            return {1} if abit else {0, 1}
        return {0, 1}
    for d in (-1, 0, 1):
        for a in (0, 1):
            for b in (0, 1):
                assert (abs(2 * d + b - a) <= 2) == (b in allowed(d, a))
