"""The warp-level view rejection of the tile kernel (kernels.cu: warp_view_reject_sphere / warp_view_reject_lean) is a shortcut in
front of Frustum::intersects_sphere (crates/bevy_camera/src/primitives.rs:255-268): it may only reject a view for a warp when the
exact test rejects that view for EVERY row of the warp.  The device code cannot run here; this restates both bounds in float32
numpy, operation for operation, and checks that property (and that the shortcut is not vacuous) on warps of rows the way the
scenes lay them out: 32 neighbours, far away from most frusta, plus adversarial cases right at a plane."""
import numpy as np
import pytest

f32 = np.float32


def exact_rejects(planes, c, r):
    """intersects_sphere, per row: True where some of the 5 planes has dot4(plane, (c, 1)) + r <= 0 (glam order)."""
    out = np.zeros(len(c), bool)
    for n in planes:
        d = (n[0] * c[:, 0] + n[2] * c[:, 2]) + (n[1] * c[:, 1] + n[3] * f32(1.0))
        out |= (d + r) <= f32(0.0)
    return out


def sphere_rejects(planes, c, r):
    """warp_view_reject_sphere for one view: a plane the warp's bounding sphere is behind."""
    x0, y0, z0 = c[0]
    mine = ((np.abs(c[:, 0] - x0) + np.abs(c[:, 1] - y0)) + np.abs(c[:, 2] - z0)) + np.abs(r)
    rmax = mine.max()
    for n in planes:
        vlen = max(f32(np.sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2])) * f32(1.000001), f32(1.0))
        reach = f32(vlen * rmax)
        d = ((n[0] * x0 + n[1] * y0) + n[2] * z0) + n[3]
        mag = ((abs(n[0] * x0) + abs(n[1] * y0)) + abs(n[2] * z0)) + (abs(n[3]) + reach)
        if (d + reach) + (f32(1e-5) * mag + f32(1e-6)) < f32(0.0):
            return True
    return False


def box_rejects(planes, c, r):
    """warp_view_reject_lean for one view: a plane the warp's bounding box (+ largest radius) is behind."""
    lo, hi, r1 = c.min(0), c.max(0), r.max()
    for n in planes:
        m = ((max(n[0] * lo[0], n[0] * hi[0]) + max(n[1] * lo[1], n[1] * hi[1])) + max(n[2] * lo[2], n[2] * hi[2])) + n[3]
        mag = ((abs(n[0]) * max(abs(lo[0]), abs(hi[0])) + abs(n[1]) * max(abs(lo[1]), abs(hi[1]))) +
               abs(n[2]) * max(abs(lo[2]), abs(hi[2]))) + (abs(n[3]) + abs(r1))
        if (m + r1) + (f32(1e-5) * mag + f32(1e-6)) < f32(0.0):
            return True
    return False


def random_frustum(rng, normalised=True):
    """Five half spaces (normal, d) of a perspective-like frustum at a random pose; optionally with un-normalised normals."""
    q = rng.normal(size=(3, 3)); q, _ = np.linalg.qr(q)
    eye = rng.uniform(-300, 300, 3)
    a, b = np.tan(rng.uniform(0.2, 0.7)), np.tan(rng.uniform(0.15, 0.5))
    local = [(1, 0, -a), (-1, 0, -a), (0, 1, -b), (0, -1, -b), (0, 0, -1)]      # L R B T near (looking down -z)
    planes = []
    for k, v in enumerate(local):
        n = q @ (np.array(v, float) / np.linalg.norm(v))
        d = -n @ eye - (0.1 if k == 4 else 0.0)
        s = 1.0 if normalised else rng.uniform(0.05, 20.0)
        planes.append(np.array([n[0] * s, n[1] * s, n[2] * s, d * s], f32))
    return planes


@pytest.mark.parametrize("normalised", [True, False])
def test_warp_shortcuts_never_reject_a_view_the_exact_test_keeps(normalised):
    rng = np.random.default_rng(7 if normalised else 8)
    hits = {"sphere": 0, "box": 0}
    trials = 0
    for _ in range(400):
        planes = random_frustum(rng, normalised)
        for spread in (0.5, 4.0, 40.0, 400.0):
            centre = rng.uniform(-500, 500, 3)
            c = (centre + rng.normal(scale=spread, size=(32, 3))).astype(f32)
            r = rng.uniform(0.0, 1.5, 32).astype(f32)
            if rng.random() < 0.1:
                r[rng.integers(32)] = f32(-0.5)          # a user-provided negative Sphere radius
            ex = exact_rejects(planes, c, r)
            trials += 1
            for name, fn in (("sphere", sphere_rejects), ("box", box_rejects)):
                if fn(planes, c, r):
                    hits[name] += 1
                    assert ex.all(), f"{name} bound rejected a warp with a row the exact test keeps"
    # the shortcut has to fire for most far-away warps, or it is worthless
    assert hits["sphere"] > trials // 3 and hits["box"] > trials // 3


def test_rows_just_inside_a_plane_are_never_rejected():
    rng = np.random.default_rng(11)
    for _ in range(300):
        planes = random_frustum(rng)
        n = planes[rng.integers(5)].astype(np.float64)
        # a tight warp whose first row touches the plane from outside by less than its radius: the exact test keeps it
        p0 = rng.uniform(-200, 200, 3)
        p0 -= n[:3] * ((n[:3] @ p0 + n[3]) / (n[:3] @ n[:3]))          # onto the plane
        r = np.full(32, 0.5, f32)
        c = (p0 - n[:3] * 0.4999 + rng.normal(scale=1e-3, size=(32, 3))).astype(f32)
        ex = exact_rejects([planes[k] for k in range(5)], c, r)
        if not ex.all():
            assert not sphere_rejects(planes, c, r)
            assert not box_rejects(planes, c, r)
