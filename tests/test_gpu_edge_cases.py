"""Edge cases the reference's code paths define, CUDA path vs oracle, bit exact."""
import math

import numpy as np
import pytest

import bevy_b200 as bb
from bevy_b200 import scenes
from bevy_b200.scenes import Scene, Camera

from parity import OracleWorld, compare_frame, run_parity

pytestmark = pytest.mark.gpu


def _random_scene(seed, n_roots=60, max_children=4, max_depth=9, n_lights=24, shuffle_entities=True, views=3):
    """Irregular forest with every per-row feature switched on somewhere."""
    rng = np.random.default_rng(seed)
    parent, depth = [], []
    for _ in range(n_roots):
        base = len(parent)
        parent.append(scenes.NO_PARENT); depth.append(0)
        frontier = [base]
        while frontier:
            nxt = []
            for p in frontier:
                if depth[p] >= max_depth or rng.random() < 0.25:
                    continue
                for _ in range(int(rng.integers(1, max_children + 1))):
                    parent.append(p); depth.append(depth[p] + 1); nxt.append(len(parent) - 1)
            frontier = nxt
    parent = np.array(parent, np.int64)
    # rows so far are in per-tree BFS-ish creation order with parent < child: keep that (topological)
    n = len(parent)
    parent = parent.astype(np.uint32)
    # a few detached rows (ChildOf pointing outside the transform hierarchy) with their own children
    for r in rng.choice(np.arange(1, n), size=5, replace=False):
        if parent[r] != scenes.NO_PARENT:
            parent[r] = bb.DETACHED
    t = rng.uniform(-40, 40, (n, 3)); t[parent != scenes.NO_PARENT] *= 0.1
    q = scenes.random_unit_quats(rng, n)
    s = rng.uniform(0.5, 1.6, (n, 3))
    s[rng.random(n) < 0.05] *= -1.0            # mirrored
    s[rng.random(n) < 0.02] = 0.0              # degenerate scale: children's GT stops changing (set_if_neq)
    trs = np.concatenate([t, q, s], 1).astype(np.float32)
    bounds = np.zeros((n, 6), np.float32)
    bounds[:, 0:3] = rng.uniform(-1, 1, (n, 3)); bounds[:, 3:6] = rng.uniform(0.1, 3.0, (n, 3))
    kind = rng.integers(0, 10, n)
    flags = np.full(n, scenes.F_INHERITED_VISIBLE, np.uint8)
    flags[kind <= 6] |= scenes.F_HAS_AABB
    flags[kind == 7] |= scenes.F_HAS_SPHERE                       # world-space sphere in bounds[0:4]
    flags[kind == 8] |= scenes.F_HAS_AABB | scenes.F_HAS_SPHERE   # Aabb takes precedence
    # kind 9: neither -> always passes the frustum stage
    flags[rng.random(n) < 0.07] &= ~np.uint8(scenes.F_INHERITED_VISIBLE)
    flags[rng.random(n) < 0.05] |= bb.F_NO_FRUSTUM_CULLING
    flags[rng.random(n) < 0.04] |= bb.F_NO_CPU_CULLING
    flags[rng.random(n) < 0.10] |= bb.F_HAS_VIS_RANGE
    cls = rng.choice([0, 1, 1, 1, 3], n).astype(np.uint8)
    layer_mask = rng.choice([1, 1, 1, 2, 3, 0, 1 << 40], n).astype(np.uint64)
    range_mask = rng.integers(0, 8, n).astype(np.uint32)
    ent = np.arange(n, dtype=np.uint64) + np.uint64(7)
    if shuffle_entities:
        ent = rng.permutation(ent) | (rng.integers(0, 3, n).astype(np.uint64) << np.uint64(32))   # generations too
    cols = (parent, trs, bounds, flags, cls)
    pos = rng.uniform(-30, 30, (n_lights, 3)).astype(np.float32)
    lrange = np.exp(rng.uniform(math.log(0.5), math.log(40.0), n_lights)).astype(np.float32)
    cols, light_row = scenes._append_lights(cols, pos, lrange)
    parent, trs, bounds, flags, cls = cols
    L = n_lights
    layer_mask = np.concatenate([layer_mask, rng.choice([1, 1, 2, 3], L).astype(np.uint64)])
    range_mask = np.concatenate([range_mask, np.zeros(L, np.uint32)])
    ent = np.concatenate([ent, np.arange(L, dtype=np.uint64) + np.uint64(1 << 20)])
    cams = []
    for k in range(views):
        qk = scenes.quat_mul(scenes.quat_axis("y", 2.1 * k), scenes.quat_axis("x", -0.2 * k))
        cams.append(Camera(gt=scenes.quat_to_gt(qk, (3.0 * k, 1.0, -2.0 * k)), quat=qk, far=120.0))
    sc = Scene(f"random_{seed}", parent, trs, bounds, flags, cls, ent, light_row, lrange, cams,
               np.nonzero(parent == scenes.NO_PARENT)[0].astype(np.uint32))
    sc.layer_mask, sc.range_mask = layer_mask, range_mask
    sc.light_layers = layer_mask[light_row].copy()
    sc.view_layers = [1, 3, 2][:views]
    sc.view_flags = [bb.VIEW_ACTIVE, bb.VIEW_ACTIVE | bb.VIEW_NO_CPU_CULLING, bb.VIEW_ACTIVE][:views]
    sc.view_range_index = [0, -1, 2][:views]
    return sc


@pytest.mark.parametrize("seed,static_opt", [(1, True), (2, False), (3, True)])
def test_random_feature_rich_scene(seed, static_opt):
    sc = _random_scene(seed)
    pipe = bb.VisibilityPipeline(sc, static_transform_optimizations=static_opt)
    world = OracleWorld(sc, static_opt)
    rng = np.random.default_rng(100 + seed)
    try:
        for f in range(5):
            if f > 0:
                scenes.advance_cameras(sc, 0.05)
                # a sparse, random set of Changed<Transform> rows (not only roots)
                rows = np.unique(rng.integers(0, sc.n, max(sc.n // 50, 1))).astype(np.uint32)
                sc.trs[rows, 0:3] += rng.uniform(-0.5, 0.5, (len(rows), 3)).astype(np.float32)
                pipe.ctx.upload_transforms_scattered(rows, sc.trs[rows])
                world.tchanged[rows] = 1
                if f == 3:       # a frame with an inactive camera: its VisibleEntities must survive untouched
                    sc.view_flags = [bb.VIEW_ACTIVE, 0, bb.VIEW_ACTIVE]
                if f == 4:
                    sc.view_flags = [bb.VIEW_ACTIVE, bb.VIEW_ACTIVE | bb.VIEW_NO_CPU_CULLING, bb.VIEW_ACTIVE]
            pipe.update_views()
            compare_frame(pipe, world, f)
    finally:
        pipe.close()


def test_static_frames_change_nothing():
    """Steady state with no input changes: no Changed<GlobalTransform>, no Changed<ViewVisibility>."""
    sc = scenes.forest(n_trees=30, levels=5, n_lights=8)
    pipe = bb.VisibilityPipeline(sc)
    world = OracleWorld(sc)
    try:
        pipe.update_views()
        compare_frame(pipe, world, 0)
        for f in (1, 2):
            pipe.update_views()          # cluster constants follow last frame's feedback, as in the reference
            s = compare_frame(pipe, world, f)
            assert s.gt_changed_count == 0 and s.vv_changed_count == 0
    finally:
        pipe.close()


def test_deep_chain_and_wide_fanout():
    """A 700-deep chain (many tiles, many passes) next to one root with 3000 children (parents in other tiles)."""
    chain = 700
    parent = [scenes.NO_PARENT] + list(range(chain - 1))
    root = len(parent)
    parent += [scenes.NO_PARENT] + [root] * 3000
    n = len(parent)
    rng = np.random.default_rng(5)
    trs = np.zeros((n, 10), np.float32)
    trs[:, 0:3] = rng.uniform(-0.2, 0.2, (n, 3)); trs[:, 3:7] = scenes.random_unit_quats(rng, n); trs[:, 7:10] = 1.0
    bounds = np.zeros((n, 6), np.float32); bounds[:, 3:6] = 0.5
    sc = Scene("chain_fanout", np.array(parent, np.uint32), trs, bounds,
               np.full(n, scenes.F_INHERITED_VISIBLE | scenes.F_HAS_AABB, np.uint8), np.ones(n, np.uint8),
               np.arange(n, dtype=np.uint64), cameras=[scenes._camera(0.3)],
               roots=np.array([0, root], np.uint32))
    run_parity(sc, frames=3, cluster=False)
    run_parity(sc, frames=2, cluster=False, static_opt=False)


def test_hierarchy_errors():
    ctx = bb.Context(8)
    try:
        with pytest.raises(bb.B200VisError) as e:
            ctx.set_topology(np.array([1, 2, 0, bb.NO_PARENT], np.uint32), np.arange(4, dtype=np.uint64))
        assert e.value.code == 4           # B200VIS_ERR_HIERARCHY_CYCLE (panic_when_hierarchy_cycle)
        with pytest.raises(bb.B200VisError) as e:
            ctx.set_topology(np.array([bb.NO_PARENT, 9], np.uint32), np.arange(2, dtype=np.uint64))
        assert e.value.code == 5
        with pytest.raises(bb.B200VisError) as e:
            ctx.set_topology(np.zeros(9, np.uint32), np.arange(9, dtype=np.uint64))
        assert e.value.code == 6
        with pytest.raises(bb.B200VisError) as e:
            ctx.run(bb.STAGE_ALL)
        assert e.value.code == 7
    finally:
        ctx.close()


def test_empty_world_and_no_lights():
    sc = Scene("empty", np.zeros(0, np.uint32), np.zeros((0, 10), np.float32), np.zeros((0, 6), np.float32),
               np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.uint64), cameras=[scenes._camera(0.0)])
    ctx = bb.Context(1, max_views=1)
    try:
        ctx.set_topology(sc.parent, sc.entity_bits)
        ctx.upload_bounds(0, sc.bounds, sc.flags, sc.class_mask)
        ctx.set_views([bb.View.make(np.tile([0, 0, 1, 1], (6, 1)))])
        ctx.run(bb.STAGE_PROPAGATE | bb.STAGE_CULL)
        assert len(ctx.download_visible(0)) == 0
    finally:
        ctx.close()


def test_recorded_frame_constants_replay_matches_live_path():
    """b200vis_record_frame_constants / use_recorded_frame_constants (the bench's device-resident replay) gives
    the same results as uploading the constants from the host."""
    sc = scenes.forest(n_trees=50, levels=6, n_lights=16)
    pipe = bb.VisibilityPipeline(sc)
    try:
        pipe.run_frame(); pipe.read_feedback()
        scenes.advance_cameras(sc, 0.02)
        pipe.update_views()
        slot = pipe.ctx.record_frame_constants()
        pipe.run_frame()
        live = [pipe.ctx.download_visible(v).copy() for v in range(4)]
        live_cl = [tuple(a.copy() for a in pipe.ctx.download_clusters(v)) for v in range(4)]
        # perturb the host-side constants, then replay the recorded frame: results must equal the live frame
        scenes.advance_cameras(sc, 0.7)
        pipe.update_views()
        pipe.ctx.use_recorded_frame_constants(slot)
        pipe.run_frame()
        for v in range(4):
            assert (pipe.ctx.download_visible(v) == live[v]).all()
            off, idx = pipe.ctx.download_clusters(v)
            assert (off == live_cl[v][0]).all() and (idx == live_cl[v][1]).all()
        pipe.ctx.use_recorded_frame_constants(None)
    finally:
        pipe.close()


def test_update_camera_and_batched_download_equal_the_piecewise_calls():
    sc = scenes.forest(n_trees=40, levels=6, n_lights=20)
    a, b = bb.VisibilityPipeline(sc), bb.VisibilityPipeline(sc)
    try:
        for f in range(3):
            scenes.advance_cameras(sc, 0.03)
            a.update_views(); b.update_views_fast()
            a.run_frame(); b.run_frame()
            sa = a.read_feedback()
            V = len(sc.cameras)
            vis = np.zeros((V, sc.n), np.uint32); off = np.zeros((V, 4097), np.uint32); idx = np.zeros((V, 1 << 16), np.uint32)
            sb = bb.FrameStats()
            b.ctx.download_frame(sb, vis, off, idx)
            for v in range(V):
                fb = b.feedback[v]
                fb.has_farthest_z, fb.farthest_z, fb.has_index_count, fb.index_count = 1, sb.cluster_farthest_z[v], 1, sb.cluster_index_count[v]
                assert sa.visible_count[v] == sb.visible_count[v] and sa.cluster_index_count[v] == sb.cluster_index_count[v]
                assert (a.ctx.download_visible(v) == vis[v, :sb.visible_count[v]]).all()
                o, i = a.ctx.download_clusters(v)
                nc = a.cluster_views[v].dims[0] * a.cluster_views[v].dims[1] * a.cluster_views[v].dims[2]
                assert tuple(a.cluster_views[v].dims) == tuple(b.cluster_views[v].dims)
                assert (o[:nc + 1] == off[v, :nc + 1]).all() and (i == idx[v, :off[v, nc]]).all()
    finally:
        a.close(); b.close()


def test_back_to_back_frames_pipelined_tail_matches_oracle():
    """Frames submitted back to back with no host synchronisation in between: the tail of frame f (visible-list
    expansion, cluster kernels on the side stream) overlaps frame f+1's tile pass.  Only the LAST frame is read
    back; it must equal the oracle stepped through the same frames.  The cluster config is feedback-free
    (constant far plane, no dynamic resizing) so no per-frame read-back is needed."""
    sc = scenes.forest(n_trees=300, levels=8, n_lights=48)
    cfg = bb.host_default_cluster_config(*sc.screen)
    cfg.far_z_mode, cfg.far_z_constant, cfg.dynamic_resizing = 1, 90.0, 0
    kw = dict(far_z_mode=1, far_z_constant=90.0, dynamic_resizing=False)
    pipe = bb.VisibilityPipeline(sc, cluster_config=cfg)
    world = OracleWorld(sc, cluster_kwargs=kw)
    try:
        frames = 7
        for f in range(frames):
            if f:
                scenes.advance_cameras(sc, 0.01)
                rows, trs = scenes.mutate_roots(sc, f)
                pipe.ctx.upload_transforms_scattered(rows, trs)
                world.tchanged[rows] = 1
            pipe.update_views()
            if f < frames - 1:
                planes = np.stack([np.ctypeslib.as_array(v.half_spaces).reshape(6, 4).copy() for v in pipe.views])
                world.frame(planes)
                pipe.run_frame()             # no download, no sync: the next frame is enqueued right behind
            else:
                compare_frame(pipe, world, f)
    finally:
        pipe.close()


def test_result_sink_matches_downloads():
    """The publish kernels write the same stats / visible rows / cluster CSR into pinned host memory that the
    download calls return."""
    torch = pytest.importorskip("torch")
    import ctypes
    sc = scenes.forest(n_trees=60, levels=6, n_lights=24)
    pipe = bb.VisibilityPipeline(sc)
    V = len(sc.cameras)
    vis = torch.zeros((V, sc.n), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    off = torch.zeros((V, 4097), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    idx = torch.zeros((V, 1 << 16), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    st_t = torch.zeros(ctypes.sizeof(bb.FrameStats), dtype=torch.uint8).pin_memory()
    st = bb.FrameStats.from_address(st_t.data_ptr())
    try:
        pipe.ctx.set_result_sink(st_t.data_ptr(), vis, off, idx)
        for f in range(3):
            scenes.advance_cameras(sc, 0.05)
            rows, trs = scenes.mutate_roots(sc, f + 1)
            pipe.ctx.upload_transforms_scattered(rows, trs)
            pipe.update_views_fast()
            pipe.run_frame()
            pipe.ctx.synchronize()
            ref = pipe.ctx.download_frame_stats()
            assert st.frame == ref.frame and st.gt_changed_count == ref.gt_changed_count and st.vv_changed_count == ref.vv_changed_count
            for v in range(V):
                assert st.visible_count[v] == ref.visible_count[v] and st.cluster_index_count[v] == ref.cluster_index_count[v]
                assert st.cluster_farthest_z[v] == ref.cluster_farthest_z[v]
                assert (vis[v, :st.visible_count[v]] == pipe.ctx.download_visible(v)).all()
                o, i = pipe.ctx.download_clusters(v)
                nc = pipe.cluster_views[v].dims[0] * pipe.cluster_views[v].dims[1] * pipe.cluster_views[v].dims[2]
                assert (off[v, :nc + 1] == o[:nc + 1]).all() and (idx[v, :off[v, nc]] == i).all()
                fb = pipe.feedback[v]
                fb.has_farthest_z, fb.farthest_z, fb.has_index_count, fb.index_count = 1, st.cluster_farthest_z[v], 1, st.cluster_index_count[v]
        pipe.ctx.set_result_sink(None, None, None, None)
    finally:
        pipe.close()


def test_step_call_equals_the_piecewise_sequence():
    """b200vis_step (upload + cameras with internal feedback + run + wait) against the explicit call sequence."""
    sc_a = scenes.forest(n_trees=40, levels=6, n_lights=20)
    sc_b = scenes.forest(n_trees=40, levels=6, n_lights=20)
    a, b = bb.VisibilityPipeline(sc_a), bb.VisibilityPipeline(sc_b)
    V = len(sc_a.cameras)
    try:
        for f in range(4):
            for sc in (sc_a, sc_b):
                scenes.advance_cameras(sc, 0.03)
            rows, trs = scenes.mutate_roots(sc_a, f + 1)
            scenes.mutate_roots(sc_b, f + 1)
            a.ctx.upload_transforms_scattered(rows, trs)
            a.update_views(); a.run_frame(); sa = a.read_feedback()
            arr = (bb.CameraDesc * V)()
            for v, cam in enumerate(sc_b.cameras):
                arr[v].global_transform[:] = cam.gt.tolist()
                arr[v].fov_y, arr[v].aspect, arr[v].near_z, arr[v].far_z = cam.fov, cam.aspect, cam.near, cam.far
                arr[v].layer_mask, arr[v].flags, arr[v].range_view_index = 1, bb.VIEW_ACTIVE, -1
            r = np.ascontiguousarray(rows, np.uint32); t_ = np.ascontiguousarray(trs, np.float32)
            b.ctx.step(len(r), r.ctypes.data, t_.ctypes.data, arr, V, b.cluster_config, wait=True)
            sb = b.ctx.download_frame_stats()
            for v in range(V):
                assert sa.visible_count[v] == sb.visible_count[v] and sa.cluster_index_count[v] == sb.cluster_index_count[v]
                assert sa.cluster_farthest_z[v] == sb.cluster_farthest_z[v]
                assert (a.ctx.download_visible(v) == b.ctx.download_visible(v)).all()
                oa, ia = a.ctx.download_clusters(v); ob, ib = b.ctx.download_clusters(v)
                assert (oa == ob).all() and (ia == ib).all()
    finally:
        a.close(); b.close()


def test_sphere_from_gt_rows_that_are_not_lights_and_changing_light_lists():
    """Rows flagged F_SPHERE_FROM_GT that were never passed to set_lights (e.g. spot lights, whose Sphere is GT-centred too,
    spot_light.rs:221) must not disturb the light snapshot the tile kernel publishes; neither may rows that stop being
    lights when the list changes.  (The snapshot is keyed on a per-row light-ordinal column, not on the bounds.)"""
    sc = scenes.forest(n_trees=30, levels=6, n_lights=24)
    L = len(sc.light_row)
    keep = np.arange(0, L, 2)                      # only every other sphere-from-GT row is a clustered light
    sc.bounds[sc.light_row, 0] = np.linspace(0.0, 1.0, L, dtype=np.float32)   # centre.x is user data now (0.0f and 1.0f included)
    all_rows, all_range = sc.light_row.copy(), sc.light_range.copy()
    sc.light_row, sc.light_range = all_rows[keep], all_range[keep]
    pipe = bb.VisibilityPipeline(sc, max_lights=L)
    world = OracleWorld(sc)
    try:
        for f in range(4):
            if f == 2:                             # the list changes: the other half becomes the light set
                other = np.arange(1, L, 2)
                sc.light_row, sc.light_range = all_rows[other], all_range[other]
                pipe.ctx.set_lights(sc.light_row, sc.light_range, None)
            if f:
                scenes.advance_cameras(sc, 0.05)
                rows, trs = scenes.mutate_roots(sc, f)
                pipe.ctx.upload_transforms_scattered(rows, trs)
                world.tchanged[rows] = 1
            pipe.update_views()
            compare_frame(pipe, world, f)
    finally:
        pipe.close()


def test_view_count_drops_and_rises_again():
    """2 -> 1 -> 2 cameras through b200vis_step (which sets the view count every frame): the per-view chunk counters of the
    view that paused must not carry counts over (visible_count / list bases would be inflated)."""
    sc = scenes.forest(n_trees=150, levels=6, n_lights=8)
    sc.cameras = sc.cameras[:2]
    pipe = bb.VisibilityPipeline(sc)
    ref = bb.VisibilityPipeline(scenes.forest(n_trees=150, levels=6, n_lights=8))     # never pauses a view
    ref.scene.cameras = ref.scene.cameras[:2]
    try:
        def cams(scene, k):
            arr = (bb.CameraDesc * k)()
            for v, cam in enumerate(scene.cameras[:k]):
                arr[v].global_transform[:] = cam.gt.tolist()
                arr[v].fov_y, arr[v].aspect, arr[v].near_z, arr[v].far_z = cam.fov, cam.aspect, cam.near, cam.far
                arr[v].layer_mask, arr[v].flags, arr[v].range_view_index = 1, bb.VIEW_ACTIVE, -1
            return arr
        for f, k in enumerate([2, 2, 1, 1, 1, 2, 2, 2]):
            for p in (pipe, ref):
                scenes.advance_cameras(p.scene, 0.04)
            pipe.ctx.step(0, 0, 0, cams(sc, k), k, None, wait=True)
            ref.ctx.step(0, 0, 0, cams(ref.scene, 2), 2, None, wait=True)
            for v in range(k):
                got, want = pipe.ctx.download_visible(v), ref.ctx.download_visible(v)
                assert len(got) == len(want) and (got == want).all(), f"frame {f} view {v}: {len(got)} vs {len(want)}"
    finally:
        pipe.close(); ref.close()


def test_result_sink_stats_through_step_with_clusters():
    """b200vis_step runs the cluster stages in a second b200vis_run call: the change counters in the sink must still be the
    CULL frame's (they were read from the next frame's already-zeroed slot)."""
    torch = pytest.importorskip("torch")
    import ctypes
    sc = scenes.forest(n_trees=60, levels=6, n_lights=24)
    pipe = bb.VisibilityPipeline(sc)
    V = len(sc.cameras)
    st_t = torch.zeros(ctypes.sizeof(bb.FrameStats), dtype=torch.uint8).pin_memory()
    st = bb.FrameStats.from_address(st_t.data_ptr())
    try:
        pipe.ctx.set_result_sink(st_t.data_ptr(), None, None, None)
        for f in range(4):
            scenes.advance_cameras(sc, 0.05)
            rows, trs = scenes.mutate_roots(sc, f + 1)
            r = np.ascontiguousarray(rows, np.uint32); t_ = np.ascontiguousarray(trs, np.float32)
            arr = (bb.CameraDesc * V)()
            for v, cam in enumerate(sc.cameras):
                arr[v].global_transform[:] = cam.gt.tolist()
                arr[v].fov_y, arr[v].aspect, arr[v].near_z, arr[v].far_z = cam.fov, cam.aspect, cam.near, cam.far
                arr[v].layer_mask, arr[v].flags, arr[v].range_view_index = 1, bb.VIEW_ACTIVE, -1
            pipe.ctx.step(len(r), r.ctypes.data, t_.ctypes.data, arr, V, pipe.cluster_config, wait=True)
            got = (st.gt_changed_count, st.vv_changed_count, st.frame)
            ref = pipe.ctx.download_frame_stats()
            assert got == (ref.gt_changed_count, ref.vv_changed_count, ref.frame), (f, got)
            _, ch = pipe.ctx.download_global_transforms(0, sc.n)
            assert got[0] == int(ch.sum()) and got[0] > 0
        pipe.ctx.set_result_sink(None, None, None, None)
    finally:
        pipe.close()


@pytest.mark.parametrize("stride", [16, 12])
def test_column_write_back_keeps_a_host_mirror_identical_to_the_device(stride):
    """b200vis_set_column_sinks + b200vis_step(WRITEBACK): the GPU writes only the CHANGED GlobalTransforms into the host
    column (glam Affine3A stride 16 or packed 12), the ViewVisibility bytes, and both change-flag bit sets.  A host mirror
    updated only through the sink must stay identical to a full download, frame after frame (static frames included)."""
    torch = pytest.importorskip("torch")
    sc = scenes.forest(n_trees=70, levels=6, n_lights=12)
    pipe = bb.VisibilityPipeline(sc)
    n, V = sc.n, len(sc.cameras)
    W = (n + 31) // 32
    gt_h = torch.zeros((n, stride), dtype=torch.float32).pin_memory().numpy()
    ident = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.float32)
    gt_h[:] = ident if stride == 12 else np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0], np.float32)
    gbits = torch.zeros(W, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    vbits = torch.zeros(W, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    vv_h = torch.zeros(n, dtype=torch.uint8).pin_memory().numpy()

    def unpack(bits):
        return np.unpackbits(bits.view(np.uint8), bitorder="little")[:n]
    try:
        pipe.ctx.set_column_sinks(gt_h, gbits, vv_h, vbits)
        for f in range(5):
            moving = f not in (2,)                          # frame 2 is static: nothing may be written
            if moving:
                scenes.advance_cameras(sc, 0.05)
                rows, trs = scenes.mutate_roots(sc, f + 1)
                if f == 3:                                  # the reference bench's pattern: only 8 roots move
                    rows, trs = rows[:8], trs[:8]
            else:
                rows, trs = np.zeros(0, np.uint32), np.zeros((0, 10), np.float32)
            r = np.ascontiguousarray(rows, np.uint32); t_ = np.ascontiguousarray(trs, np.float32)
            arr = (bb.CameraDesc * V)()
            for v, cam in enumerate(sc.cameras):
                arr[v].global_transform[:] = cam.gt.tolist()
                arr[v].fov_y, arr[v].aspect, arr[v].near_z, arr[v].far_z = cam.fov, cam.aspect, cam.near, cam.far
                arr[v].layer_mask, arr[v].flags, arr[v].range_view_index = 1, bb.VIEW_ACTIVE, -1
            before = gt_h.copy()
            pipe.ctx.step(len(r), r.ctypes.data if len(r) else 0, t_.ctypes.data if len(r) else 0, arr, V, pipe.cluster_config,
                          wait=True, writeback=True)
            pipe.ctx.synchronize()
            gt, ch = pipe.ctx.download_global_transforms(0, n, stride=stride)
            vv, vch = pipe.ctx.download_view_visibility(0, n)
            assert (unpack(gbits) == ch).all() and (unpack(vbits) == vch).all(), f
            assert (gt_h.view(np.uint32) == gt.view(np.uint32)).all(), f"frame {f}: host mirror differs from the device column"
            assert (gt_h[ch == 0].view(np.uint32) == before[ch == 0].view(np.uint32)).all()
            assert (vv_h == vv).all()
            if f == 2:
                assert ch.sum() == 0
            if f == 3:
                assert 0 < ch.sum() <= 8 * 63
        pipe.ctx.set_column_sinks()
    finally:
        pipe.close()


def test_visible_entities_per_visibility_class():
    """VisibleEntities::entities is one sorted Vec per VisibilityClass; an entity with k classes is pushed k times
    (visibility/mod.rs:344-347, 852-857).  The device returns each view's sorted list plus the class mask of every entry; the
    split the shim performs must equal the oracle's push-per-class + sort-per-class restatement."""
    import oracle as orc
    sc = scenes.forest(n_trees=90, levels=6, n_lights=10)
    rng = np.random.default_rng(5)
    sc.class_mask = rng.choice(np.array([0, 1, 2, 3, 6, 0x81, 0xFF], np.uint8), sc.n).astype(np.uint8)
    sc.entity_bits = rng.permutation(sc.n).astype(np.uint64) + np.uint64(7)          # rows are NOT in Entity order
    pipe = bb.VisibilityPipeline(sc)
    world = OracleWorld(sc)
    try:
        for f in range(3):
            if f:
                scenes.advance_cameras(sc, 0.08)
                rows, trs = scenes.mutate_roots(sc, f)
                pipe.ctx.upload_transforms_scattered(rows, trs)
                world.tchanged[rows] = 1
            pipe.update_views()
            compare_frame(pipe, world, f)
            for v in range(len(sc.cameras)):
                got = pipe.ctx.download_visible_by_class(v)
                want = orc.visible_entities_by_class(world.last_lists[v], sc.class_mask, sc.entity_bits)
                assert sorted(got) == sorted(want), (f, v, sorted(got), sorted(want))
                for k in want:
                    assert len(got[k]) == len(want[k]) and (got[k] == want[k]).all(), f"frame {f} view {v} class {k}"
                assert sum(len(x) for x in want.values()) >= len(world.last_lists[v])     # multi-class rows are pushed more than once
    finally:
        pipe.close()


def test_render_layers_beyond_the_first_64():
    """RenderLayers is a SmallVec of 64-bit blocks (render_layers.rs:20-23): an entity on layer 70 and a camera on layers
    {3, 70} intersect through block 1 although their first blocks do not (intersects(), :121-135)."""
    import oracle as orc
    sc = scenes.forest(n_trees=60, levels=6, n_lights=8)
    rng = np.random.default_rng(3)
    n, V = sc.n, len(sc.cameras)
    sc.layer_mask = rng.choice(np.array([0, 1, 2, 8], np.uint64), n).astype(np.uint64)      # block 0 (0: none of the first 64 layers)
    ext = np.zeros((n, 3), np.uint64)
    pick = rng.random(n) < 0.5
    ext[pick, rng.integers(0, 3, pick.sum())] = np.uint64(1) << rng.integers(0, 64, pick.sum()).astype(np.uint64)
    sc.view_layers = [1, 0, 2, 8][:V]
    view_ext = np.zeros((V, 3), np.uint64)
    view_ext[1 % V] = [0xFFFFFFFFFFFFFFFF, 0, 0]
    view_ext[2 % V, 2] = 0xFFFFFFFF00000000
    pipe = bb.VisibilityPipeline(sc)
    world = OracleWorld(sc)
    orc.set_render_layers_ext(ext, view_ext)
    try:
        pipe.ctx.upload_render_layers_ext(0, ext)
        for v in range(V):
            pipe.ctx.set_view_render_layers_ext(v, view_ext[v])
        for f in range(3):
            if f:
                scenes.advance_cameras(sc, 0.1)
                rows, trs = scenes.mutate_roots(sc, f)
                pipe.ctx.upload_transforms_scattered(rows, trs)
                world.tchanged[rows] = 1
            pipe.update_views()
            compare_frame(pipe, world, f)
        only_ext = (sc.layer_mask[world.last_lists[1 % V]] & np.uint64(sc.view_layers[1 % V])) == 0
        assert only_ext.any()          # some rows are listed through a block beyond the first
    finally:
        orc.set_render_layers_ext(None, None)
        pipe.close()
