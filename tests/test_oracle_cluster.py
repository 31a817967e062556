"""The reference has NO test for cluster membership, so the oracle's restatement of the
iterative sphere refinement is cross-checked by geometry: every cluster that truly intersects the light's
view-space sphere must be listed (no false negatives), and everything listed must lie inside the
conservative cluster range of the light's view-space AABB."""
import math

import numpy as np

import oracle as orc
from bevy_b200 import scenes


def _setup(seed, n_lights=40):
    rng = np.random.default_rng(seed)
    q = scenes.quat_mul(scenes.quat_axis("y", rng.uniform(0, 6.28)), scenes.quat_axis("x", rng.uniform(-0.5, 0.5)))
    cam = scenes.quat_to_gt(q, rng.uniform(-5, 5, 3))
    cfv = orc.perspective(math.pi / 4, 16 / 9, 0.1)
    fr = orc.compute_frustum(cfv, cam, 1000.0)
    inv = orc.affine_inverse(cam)
    # lights in front of the camera (view space), then to world space
    pv = np.stack([rng.uniform(-30, 30, n_lights), rng.uniform(-15, 15, n_lights), -rng.uniform(1, 120, n_lights)], 1)
    R = cam[:9].reshape(3, 3).T.astype(np.float64)      # columns x,y,z
    pw = (pv @ R.T + cam[9:12]).astype(np.float32)
    rng_l = np.exp(rng.uniform(math.log(0.5), math.log(25.0), n_lights)).astype(np.float32)
    lights = np.concatenate([pw, rng_l[:, None]], 1).astype(np.float32)
    vin = orc.default_cluster_view_in(cam, cfv, fr, last_farthest_z=150.0)
    return vin, lights, inv, cfv


def test_refinement_has_no_false_negatives_and_stays_in_aabb_range():
    for seed in range(6):
        vin, lights, inv, cfv = _setup(seed)
        out, offsets, idx, planes = orc.assign_lights_to_clusters(vin, lights, want_planes=True)
        dx, dy, dz = out.dims
        xp, yp, zp = [p.astype(np.float64) for p in planes]
        member = np.zeros((dx * dy * dz, len(lights)), bool)
        for c in range(dx * dy * dz):
            member[c, idx[offsets[c]:offsets[c + 1]]] = True
        assert out.total_index_count == member.sum()
        m = inv.astype(np.float64)
        Rm = np.stack([m[0:3], m[3:6], m[6:9]], 1)
        rngs = np.random.default_rng(seed + 100)
        for li, (x, y, z, r) in enumerate(lights.astype(np.float64)):
            cv = Rm @ np.array([x, y, z]) + m[9:12]
            # sample points inside the sphere; the cluster containing each sample must list the light
            pts = rngs.normal(size=(300, 3)); pts /= np.linalg.norm(pts, axis=1, keepdims=True)
            pts = cv + pts * (rngs.uniform(0, 1, (300, 1)) ** (1 / 3)) * r * 0.999
            for p in pts:
                if p[2] >= -1e-3:
                    continue
                # cluster coordinates of the point from the plane tables (inside-facing normals)
                ix = np.nonzero((xp[:, [0, 2]] @ p[[0, 2]]) >= 0)[0]
                iy = np.nonzero((yp[:, [1, 2]] @ p[[1, 2]]) >= 0)[0]
                if len(ix) == 0 or len(iy) == 0 or len(ix) > dx or len(iy) > dy:
                    continue                               # outside the screen
                cx, cy = ix.max(), iy.max()
                depth = -p[2]
                zs = -zp[:, 3] * np.sign(zp[:, 2]) if False else np.array([(-zp[k, 3] / zp[k, 2]) for k in range(dz + 1)])
                below = np.nonzero(-zs <= depth)[0]
                if len(below) == 0 or below.max() >= dz:
                    continue                               # beyond the far slice
                cz = below.max()
                c = (cy * dx + cx) * dz + cz
                assert member[c, li], f"seed {seed}: light {li} misses cluster {(cx, cy, cz)}"


def test_total_index_count_and_farthest_z_are_consistent():
    vin, lights, inv, cfv = _setup(11)
    out, offsets, idx, _ = orc.assign_lights_to_clusters(vin, lights)
    assert offsets[-1] == out.total_index_count == len(idx)
    m = inv.astype(np.float64)
    far = max(0.0, max(-(m[2] * x + m[5] * y + m[8] * z + m[11]) + r for x, y, z, r in lights.astype(np.float64)))
    assert abs(out.farthest_z - far) < 1e-3 * max(1.0, far)


def test_lights_out_of_frustum_or_wrong_layer_are_skipped():
    vin, lights, _, _ = _setup(5, 10)
    out_all, _, idx_all, _ = orc.assign_lights_to_clusters(vin, lights)
    layers = np.ones(len(lights), np.uint64); layers[::2] = 2      # view is on layer 0 only
    out, offsets, idx, _ = orc.assign_lights_to_clusters(vin, lights, layers)
    assert set(idx.tolist()) <= set(range(1, len(lights), 2))
    behind = lights.copy(); behind[:, :3] = 1e6
    out_b, off_b, idx_b, _ = orc.assign_lights_to_clusters(vin, behind)
    assert out_b.total_index_count == 0 and out_b.farthest_z == 0.0
