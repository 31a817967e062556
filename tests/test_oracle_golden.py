"""Pins the CPU oracle against every known-answer vector the reference's own
tests hold for this path.  The vectors are transcribed
as DATA from the cited reference tests; the code under test is oracle/."""
import math

import numpy as np
import pytest

import oracle as orc

TAU = np.float32(2.0 * math.pi)


# --- crates/bevy_camera/src/primitives.rs:462-611 ------------------------------
BIG_FRUSTUM = [(-0.9701, -0.2425, -0.0000, 7.7611), (-0.0000, 1.0000, -0.0000, 4.0000),
               (-0.0000, -0.2425, -0.9701, 2.9104), (-0.0000, -1.0000, -0.0000, 4.0000),
               (-0.0000, -0.2425, 0.9701, 2.9104), (0.9701, -0.2425, -0.0000, -1.9403)]
FRUSTUM = [(-0.9701, -0.2425, -0.0000, 0.7276), (-0.0000, 1.0000, -0.0000, 1.0000),
           (-0.0000, -0.2425, -0.9701, 0.7276), (-0.0000, -1.0000, -0.0000, 1.0000),
           (-0.0000, -0.2425, 0.9701, 0.7276), (0.9701, -0.2425, -0.0000, 0.7276)]
LONG_FRUSTUM = [(-0.9998, -0.0222, -0.0000, -1.9543), (-0.0000, 1.0000, -0.0000, 45.1249),
                (-0.0000, -0.0168, -0.9999, 2.2718), (-0.0000, -1.0000, -0.0000, 45.1249),
                (-0.0000, -0.0168, 0.9999, 2.2718), (0.9998, -0.0222, -0.0000, 7.9528)]

SPHERE_CASES = [
    # (frustum, center, radius, expected)  -- primitives.rs test name
    (BIG_FRUSTUM, (0.9167, 0.0, 0.0), 0.75, False),      # intersects_sphere_big_frustum_outside
    (BIG_FRUSTUM, (7.9288, 0.0, 2.9728), 2.0, True),     # intersects_sphere_big_frustum_intersect
    (FRUSTUM, (0.0, 0.0, 0.0), 3.0, True),               # intersects_sphere_frustum_surrounding
    (FRUSTUM, (0.0, 0.0, 0.0), 0.7, True),               # intersects_sphere_frustum_contained
    (FRUSTUM, (0.0, 0.0, 0.9695), 0.7, True),            # intersects_sphere_frustum_intersects_plane
    (FRUSTUM, (1.2037, 0.0, 0.9695), 0.7, True),         # ..._intersects_2_planes
    (FRUSTUM, (1.2037, -1.0988, 0.9695), 0.7, True),     # ..._intersects_3_planes
    (FRUSTUM, (-1.7020, 0.0, 0.0), 0.7, False),          # ..._dodges_1_plane
    (LONG_FRUSTUM, (-4.4889, 46.9021, 0.0), 0.75, False),  # intersects_sphere_long_frustum_outside
    (LONG_FRUSTUM, (-4.9957, 0.0, -0.7396), 4.4094, True),  # intersects_sphere_long_frustum_intersect
]


@pytest.mark.parametrize("frustum,center,radius,expected", SPHERE_CASES)
def test_intersects_sphere_known_answers(frustum, center, radius, expected):
    planes = orc.frustum_from_half_spaces(frustum)
    assert orc.intersects_sphere(planes, center, radius, True) is expected


def _affine_rot_trans(quat, t):
    return orc.affine_from_trs(np.array([*t, *quat, 1, 1, 1], np.float32))


def _quat_axis(axis, angle):
    # glam Quat::from_rotation_{x,y,z}: (sin(a/2) on the axis, cos(a/2))
    s, c = math.sin(angle * 0.5), math.cos(angle * 0.5)
    q = [0.0, 0.0, 0.0, c]
    q["xyz".index(axis)] = s
    return np.array(q, np.float32)


# --- primitives.rs:614-685 -----------------------------------------------------
def test_sphere_intersects_obb_cases():
    I = orc.IDENTITY_GT
    assert orc.sphere_intersects_obb((0, 0, 0), 1.0, (0, 0, 0), (0.5, 0.5, 0.5), I)      # identical_center
    assert orc.sphere_intersects_obb((1, 0, 0), 0.0, (0, 0, 0), (1, 0, 0), I)            # at_edge
    assert orc.sphere_intersects_obb((0, 0, 0), 10.0, (1, 1, 1), (0, 0, 0), I)           # zero_extents_inside
    t = _affine_rot_trans(_quat_axis("y", math.pi), (5.0, 0.0, 0.0))                      # rotated_zeros
    assert orc.sphere_intersects_obb((5, 0, 0), 1.0, (0, 0, 0), (0, 0, 0), t)


# --- primitives.rs:712-799 (contains_aabb with a real PerspectiveProjection) ----
def _contains_aabb_test_frustum():
    cfv = orc.perspective(np.float32(math.radians(90.0)), 1.0, 1.0)
    cam = orc.IDENTITY_GT.copy(); cam[9:12] = (2.0, 2.0, 0.0)
    return orc.compute_frustum(cfv, cam, 100.0)


def _contains_aabb_test_frustum_with_rotation():
    f32 = np.float32
    half_extent_world = f32(math.sqrt(f32((49.5 * 49.5) * 0.5))) + f32(math.sqrt(f32(0.5)))
    near = f32(50.5) - half_extent_world
    far = near + f32(2.0) * half_extent_world
    fov = f32(2.0) * f32(math.atan(half_extent_world / near))
    cfv = orc.perspective(fov, 1.0, near)
    return orc.compute_frustum(cfv, orc.IDENTITY_GT, far)


def _translation(t):
    g = orc.IDENTITY_GT.copy(); g[9:12] = t; return g


def test_contains_aabb_known_answers():
    fr = _contains_aabb_test_frustum()
    assert orc.contains_aabb(fr, (0, 0, 0), (0.99, 0.99, 49.49), _translation((2, 2, -50.5)))       # aabb_inside_frustum
    assert not orc.contains_aabb(fr, (0, 0, 0), (0.99, 0.99, 49.6), _translation((2, 2, -50.5)))    # aabb_intersect_frustum
    assert not orc.contains_aabb(fr, (0, 0, 0), (0.99, 0.99, 0.99), _translation((0, 0, 49.6)))     # aabb_outside_frustum
    fr = _contains_aabb_test_frustum_with_rotation()
    model = _affine_rot_trans(_quat_axis("x", math.pi / 4.0), (0.0, 0.0, -50.5))
    assert orc.contains_aabb(fr, (0, 0, 0), (0.99, 0.99, 49.49), model)                            # aabb_inside_frustum_rotation
    assert not orc.contains_aabb(fr, (0, 0, 0), (0.99, 0.99, 49.6), model)                         # aabb_intersect_frustum_rotation


# --- primitives.rs:834-857 (identity-path equivalence) --------------------------
def test_intersects_obb_identity_matches_standard():
    aabbs = [((0, 0, 0), (0.5, 0.5, 0.5)), ((1.0, 0.0, 0.5), (0.9, 0.9, 0.9)), ((100, 100, 100), (1, 1, 1))]
    for fr in (FRUSTUM, LONG_FRUSTUM, BIG_FRUSTUM):
        planes = orc.frustum_from_half_spaces(fr)
        for c, h in aabbs:
            assert orc.intersects_obb(planes, c, h, orc.IDENTITY_GT, True, True) == \
                orc.intersects_obb_identity(planes, c, h)


# --- benches/benches/bevy_camera/primitives.rs:12-57: both asserted true ---------
def test_bench_fixture_obb_true():
    planes = orc.frustum_from_half_spaces(FRUSTUM)
    assert orc.intersects_obb(planes, (0, 0, 0), (0.5, 0.5, 0.5), orc.IDENTITY_GT, True, True)


# --- crates/bevy_math/src/primitives/half_space.rs:53-57 -------------------------
def test_half_space_new_normalises():
    hs = orc.half_space_new((0.0, 3.0, 4.0, 10.0))
    np.testing.assert_allclose(hs, [0.0, 0.6, 0.8, 2.0], rtol=1e-6)
    hs = orc.half_space_new((0.0, 0.0, 0.0, np.inf))   # INACTIVE_HALF_SPACE (view_frustum.rs:38)
    assert np.isnan(hs[:3]).all()                       # 0 * inf


# --- crates/bevy_transform/src/helper.rs:98-146 ----------------------------------
def _chain_case(transforms):
    n = len(transforms)
    parent = np.array([orc.NO_PARENT] + list(range(n - 1)), np.uint32)
    trs = np.array(transforms, np.float32)
    gt = np.tile(orc.IDENTITY_GT, (n, 1))
    rc, changed = orc.propagate(parent, trs, gt, np.ones(n, np.uint8), static_opt=True)
    assert rc == 0
    # TransformHelper::compute_global_transform: fold from the root down
    acc = orc.affine_from_trs(trs[0])
    for i in range(1, n):
        acc = orc.affine_mul(acc, orc.affine_from_trs(trs[i]))
    return gt[-1], acc


def test_match_transform_propagation_systems():
    s = lambda v: (v, v, v)
    t0 = [1, 0, 0, *_quat_axis("y", TAU / 4), *s(2.0)]
    t1 = [0, 1, 0, *_quat_axis("z", TAU / 3), *s(1.5)]
    t2 = [0, 0, 1, *_quat_axis("x", TAU / 2), *s(0.3)]
    leaf, helper = _chain_case([t0])
    np.testing.assert_allclose(leaf, helper, atol=1.19e-7)   # approx::assert_abs_diff_eq! default eps = f32::EPSILON
    leaf, helper = _chain_case([t0, t1, t2])
    np.testing.assert_allclose(leaf, helper, atol=1.19e-7)
    # independent float64 cross-check of the composed matrix
    def mat(t):
        tr, q, sc = np.array(t[:3], float), np.array(t[3:7], float), np.array(t[7:], float)
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        M = np.eye(4); M[:3, :3] = R * sc; M[:3, 3] = tr
        return M
    M = mat(t0) @ mat(t1) @ mat(t2)
    got = np.eye(4); got[:3, 0] = leaf[0:3]; got[:3, 1] = leaf[3:6]; got[:3, 2] = leaf[6:9]; got[:3, 3] = leaf[9:12]
    np.testing.assert_allclose(got, M, atol=2e-6)


def _t(x, y, z):
    return [x, y, z, 0, 0, 0, 1, 1, 1, 1]


# --- crates/bevy_transform/src/systems.rs:889-925 (did_propagate) ------------------
def test_did_propagate():
    parent = np.array([orc.NO_PARENT, 0, 0], np.uint32)
    trs = np.array([_t(1, 0, 0), _t(0, 2, 0), _t(0, 0, 3)], np.float32)
    gt = np.tile(orc.IDENTITY_GT, (3, 1))
    rc, _ = orc.propagate(parent, trs, gt, np.ones(3, np.uint8))
    assert rc == 0
    assert tuple(gt[1, 9:12]) == (1.0, 2.0, 0.0)     # exact equality in the reference test
    assert tuple(gt[2, 9:12]) == (1.0, 0.0, 3.0)


# --- systems.rs:827-886 (correct_parent_removed): orphan keeps its own transform ---
def test_correct_parent_removed():
    # root(1,0,0)+offset -> parent(0,2,0)... the reference's scenario: after the parent is
    # removed, the child becomes a root whose GT = its own Transform.
    parent = np.array([orc.NO_PARENT, 0, 1], np.uint32)
    trs = np.array([_t(3, 0, 0), _t(0, 5, 0), _t(0, 0, 7)], np.float32)
    gt = np.tile(orc.IDENTITY_GT, (3, 1))
    orc.propagate(parent, trs, gt, np.ones(3, np.uint8))
    assert tuple(gt[2, 9:12]) == (3.0, 5.0, 7.0)
    # remove ChildOf from row 1: it becomes a root (with child 2); orphaned => tchanged
    parent = np.array([orc.NO_PARENT, orc.NO_PARENT, 1], np.uint32)
    rc, changed = orc.propagate(parent, trs, gt, np.array([0, 1, 0], np.uint8))
    assert rc == 0
    assert tuple(gt[1, 9:12]) == (0.0, 5.0, 0.0)
    assert tuple(gt[2, 9:12]) == (0.0, 5.0, 7.0)
    assert changed.tolist() == [0, 1, 1]


# --- systems.rs:1048-1097 (correct_transforms_when_no_children) --------------------
def test_correct_transforms_when_no_children():
    # translation-only chain; exact equality
    parent = np.array([orc.NO_PARENT, 0, 1], np.uint32)
    trs = np.array([_t(1, 0, 0), _t(0, 1, 0), _t(0, 0, 1)], np.float32)
    gt = np.tile(orc.IDENTITY_GT, (3, 1))
    orc.propagate(parent, trs, gt, np.ones(3, np.uint8))
    assert tuple(gt[2, 9:12]) == (1.0, 1.0, 1.0)


# --- systems.rs:1101-1164 (panic_when_hierarchy_cycle) -> error code ----------------
def test_hierarchy_cycle_is_an_error():
    parent = np.array([1, 2, 0, orc.NO_PARENT], np.uint32)
    gt = np.tile(orc.IDENTITY_GT, (4, 1))
    rc, _ = orc.propagate(parent, np.tile(_t(0, 0, 0), (4, 1)).astype(np.float32), gt, np.ones(4, np.uint8))
    assert rc == -2


# --- systems.rs:1167-1221 (global_transform_should_not_be_overwritten_after_reparenting)
def test_reparenting_keeps_externally_written_gt_of_flat_entity():
    # A flat entity whose Transform did not change keeps whatever GT it has.
    parent = np.array([orc.NO_PARENT], np.uint32)
    gt = np.array([[1, 0, 0, 0, 1, 0, 0, 0, 1, 9, 9, 9]], np.float32)
    rc, changed = orc.propagate(parent, np.array([_t(1, 2, 3)], np.float32), gt, np.zeros(1, np.uint8))
    assert changed.tolist() == [0] and tuple(gt[0, 9:12]) == (9.0, 9.0, 9.0)


# --- crates/bevy_camera/src/visibility/mod.rs:1314-1448 (view_visibility_lifecycle) --
def test_view_visibility_lifecycle():
    """The reference marks the entity visible by calling set_visible() from a system in
    CheckVisibility; here the same effect comes from one always-true view (no bounds =>
    visible) whose `active` flag plays ManualMark."""
    gt = orc.IDENTITY_GT[None].copy()
    bounds = np.zeros((1, 6), np.float32)
    flags = np.array([orc.F_INHERITED_VISIBLE], np.uint8)
    planes = orc.frustum_from_half_spaces(FRUSTUM)[None]
    vv = np.array([0], np.uint8)          # ViewVisibility::HIDDEN
    ent = np.array([1], np.uint64)
    cls = np.array([1], np.uint8)

    def frame(mark):
        ch, _ = orc.cull(gt, bounds, flags, cls, ent, vv, planes,
                         view_flags=np.array([orc.VIEW_ACTIVE if mark else 0], np.uint8))
        return bool(vv[0] & 1), bool(ch[0])

    assert frame(False) == (False, False)   # Frame 1: hidden, not changed
    assert frame(True) == (True, True)      # Frame 2: visible, changed
    assert frame(True) == (True, False)     # Frame 3: still visible, NOT changed
    assert vv[0] == 0b11
    assert frame(False) == (False, True)    # Frame 4: hidden, changed
    assert frame(False) == (False, False)   # Frame 5: hidden, NOT changed
    assert vv[0] == 0


# --- crates/bevy_light/src/cluster/test.rs:5-54 --------------------------------------
def _check_tiling(w, h):
    dims = orc.cluster_dimensions_for_screen_size(3, (0, 0, 0), 4096, 24, w, h)
    tile, cd = orc.clusters_update(w, h, dims)
    assert tile[0] * cd[0] >= w and tile[1] * cd[1] >= h
    assert tile[0] * (cd[0] - 1) < w and tile[1] * (cd[1] - 1) < h
    assert (tile[0] - 1) * cd[0] < w and (tile[1] - 1) * cd[1] < h
    assert cd[0] <= w and cd[1] <= h
    assert cd[0] * cd[1] * cd[2] <= 4096


def test_default_cluster_setup_small_screensizes():
    for x in range(1, 100):
        for y in range(1, 100):
            _check_tiling(x, y)


def test_default_cluster_setup_small_x():
    for x in range(1, 10):
        for y in range(1, 5000, 7):     # strided: the reference sweeps every y; same invariants
            _check_tiling(x, y)
            _check_tiling(y, x)


def test_default_1080p_grid_is_17x9x24():
    # 1920x1080 default => 17x9x24 = 3672 clusters
    dims = orc.cluster_dimensions_for_screen_size(3, (0, 0, 0), 4096, 24, 1920, 1080)
    tile, cd = orc.clusters_update(1920, 1080, dims)
    assert cd == (17, 9, 24)
