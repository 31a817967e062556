"""SURVEY.md 8(f) rows on the device, against the oracle restatements (bevy_oracle_next.c), bit exact."""
import numpy as np
import pytest
import torch

import bevy_b200 as bb
from bevy_b200 import scenes

from parity import OracleWorld, compare_frame, run_parity
from test_gpu_edge_cases import _random_scene

pytestmark = pytest.mark.gpu


# ---- N1: render-world visible-entity diff ------------------------------------------------------------
def test_n1_visible_diff_forest_moving_cameras():
    run_parity(scenes.forest(n_trees=60, levels=6, n_lights=8), frames=5, visible_diff=True)


@pytest.mark.parametrize("seed", [2, 3])
def test_n1_visible_diff_random_scene_with_inactive_view_and_shuffled_entities(seed):
    sc = _random_scene(seed)
    pipe = bb.VisibilityPipeline(sc)
    world = OracleWorld(sc, True)
    pipe.enable_visible_diff()
    rng = np.random.default_rng(seed)
    try:
        for f in range(6):
            if f > 0:
                scenes.advance_cameras(sc, 0.08)
                rows = np.unique(rng.integers(0, sc.n, max(sc.n // 30, 1))).astype(np.uint32)
                sc.trs[rows, 0:3] += rng.uniform(-1.5, 1.5, (len(rows), 3)).astype(np.float32)
                pipe.ctx.upload_transforms_scattered(rows, sc.trs[rows])
                world.tchanged[rows] = 1
                sc.view_flags = [bb.VIEW_ACTIVE, 0 if f in (2, 3) else bb.VIEW_ACTIVE, bb.VIEW_ACTIVE]
            pipe.update_views()
            compare_frame(pipe, world, f)
    finally:
        pipe.close()


def test_n1_diff_sink_and_reset_on_enable():
    sc = scenes.forest(n_trees=40, levels=6, n_lights=4)
    pipe = bb.VisibilityPipeline(sc)
    V = pipe.ctx.max_views
    cap = sc.n
    rows = torch.zeros((2, V, cap), dtype=torch.int32).pin_memory()
    counts = torch.zeros((V, 2), dtype=torch.int32).pin_memory()
    try:
        pipe.enable_visible_diff()
        pipe.ctx.set_visible_diff_sink(rows.numpy().view(np.uint32), counts.numpy().view(np.uint32))
        prev = [np.zeros(0, np.uint32) for _ in sc.cameras]
        for f in range(4):
            if f:
                scenes.advance_cameras(sc, 0.1)
            pipe.update_views()
            pipe.run_frame()
            pipe.ctx.synchronize()
            for v in range(len(sc.cameras)):
                cur = pipe.ctx.download_visible(v)
                na, nr = counts[v, 0].item(), counts[v, 1].item()
                assert np.array_equal(rows[0, v, :na].numpy().view(np.uint32), np.setdiff1d(cur, prev[v]))
                assert np.array_equal(rows[1, v, :nr].numpy().view(np.uint32), np.setdiff1d(prev[v], cur))
                if f == 0:
                    assert nr == 0 and na == len(cur)
                prev[v] = cur
        # re-enabling resets the old list: everything visible is reported as added again
        pipe.enable_visible_diff(False)
        pipe.enable_visible_diff(True)
        pipe.run_frame()
        a, r = pipe.ctx.download_visible_diff(0)
        assert len(r) == 0 and np.array_equal(a, pipe.ctx.download_visible(0))
    finally:
        pipe.ctx.set_visible_diff_sink(None, None)
        pipe.close()


# ---- N2: Clusters -> ViewClusterBindings wire format -----------------------------------------------------
def _big_light_scene(n_lights, light_range):
    sc = scenes.forest(n_trees=30, levels=5, n_lights=n_lights)
    sc.light_range[:] = light_range
    sc.bounds[sc.light_row, 3] = light_range
    return sc


@pytest.mark.parametrize("mode,n_lights,light_range,no_resize", [
    (1, 64, 12.0, False),     # storage buffers
    (2, 64, 6.0, False),      # uniform buffers, fits
    (2, 200, 45.0, True),     # uniform buffers, more than MAX_INDICES index slots: the record loop breaks
    (1, 200, 45.0, True),
])
def test_n2_cluster_bindings_match_the_reference_packing(mode, n_lights, light_range, no_resize):
    import oracle as orc
    from bevy_b200 import abi
    sc = _big_light_scene(n_lights, light_range)
    cfg = abi.host_default_cluster_config(*sc.screen)
    if no_resize:
        cfg.dynamic_resizing = 0
        cfg.view_cluster_bindings_max_indices = 1 << 22
    pipe = bb.VisibilityPipeline(sc, cluster_config=cfg, max_cluster_indices=1 << 20)
    rng = np.random.default_rng(5)
    gmap = rng.permutation(n_lights).astype(np.uint32) if mode == 1 else None   # uniform mode: 8-bit slots, ids < 256
    try:
        pipe.ctx.set_cluster_bindings(mode, gmap)
        saw_overflow = False
        for f in range(3):
            if f:
                scenes.advance_cameras(sc, 0.2)
            pipe.update_views()
            pipe.run_frame()
            pipe.read_feedback()
            for v in range(len(sc.cameras)):
                offsets, idx = pipe.ctx.download_clusters(v)
                cv = pipe.cluster_views[v]
                nc = cv.dims[0] * cv.dims[1] * cv.dims[2]
                w_oc, w_il, w_no, w_ni = orc.cluster_bindings(offsets[:nc + 1], idx, gmap, storage=(mode == 1))
                g_oc, g_il, g_no, g_ni = pipe.ctx.download_cluster_bindings(v)
                assert (g_no, g_ni) == (w_no, w_ni), f"frame {f} view {v}: n_offsets/n_indices {(g_no, g_ni)} vs {(w_no, w_ni)}"
                assert np.array_equal(g_oc, w_oc), f"frame {f} view {v}: offsets_and_counts differ"
                assert np.array_equal(g_il, w_il), f"frame {f} view {v}: index lists differ"
                saw_overflow |= len(idx) > 16384
        assert saw_overflow == no_resize
    finally:
        pipe.close()


# ---- N4a: check_visibility_ranges inside the cull phase -----------------------------------------------------
@pytest.mark.parametrize("seed", [4, 5])
def test_n4_visibility_ranges_computed_on_device(seed):
    sc = _random_scene(seed)
    rng = np.random.default_rng(seed)
    n = sc.n
    # every row gets VisibilityRange parameters; only rows flagged F_HAS_VIS_RANGE are in the query
    start = rng.uniform(0, 60, n).astype(np.float32)
    sc.range_se = np.stack([start, start + rng.uniform(0, 120, n).astype(np.float32)], 1)
    sc.range_use_aabb = rng.integers(0, 2, n).astype(np.uint8)
    sc.flags = sc.flags | (rng.random(n) < 0.5).astype(np.uint8) * bb.F_HAS_VIS_RANGE
    extra = rng.uniform(-60, 60, (33, 3)).astype(np.float32)     # 36 range views: only the first 32 count

    def view_positions():
        return np.concatenate([np.stack([np.asarray(c.gt, np.float32)[9:12] for c in sc.cameras]), extra])

    sc.view_range_index = np.arange(len(sc.cameras), dtype=np.int8)
    sc.range_view_pos = view_positions()
    sc.range_mask = np.zeros(n, np.uint32)
    pipe = bb.VisibilityPipeline(sc)
    world = OracleWorld(sc, True)
    try:
        pipe.ctx.upload_visibility_ranges(0, sc.range_se, sc.range_use_aabb)
        for f in range(4):
            if f:
                scenes.advance_cameras(sc, 0.1)
                for c in sc.cameras:                       # move the cameras too: distances change
                    c.gt = np.asarray(c.gt, np.float32).copy(); c.gt[9:12] += rng.uniform(-8, 8, 3).astype(np.float32)
                rows = np.unique(rng.integers(0, n, n // 20)).astype(np.uint32)
                sc.trs[rows, 0:3] += rng.uniform(-3, 3, (len(rows), 3)).astype(np.float32)
                pipe.ctx.upload_transforms_scattered(rows, sc.trs[rows])
                world.tchanged[rows] = 1
            sc.range_view_pos = view_positions()
            pipe.ctx.set_visibility_range_views(sc.range_view_pos)
            pipe.update_views()
            compare_frame(pipe, world, f)
        assert sc.range_mask.any() and (sc.range_mask[(sc.flags & bb.F_HAS_VIS_RANGE) != 0] == 0).any()
    finally:
        pipe.close()


# ---- N4b: visibility_propagate_system ------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [6, 7])
def test_n4_visibility_propagation_matches_the_change_driven_system(seed):
    import oracle as orc
    sc = _random_scene(seed, n_roots=80, max_depth=12)
    rng = np.random.default_rng(seed)
    n = sc.n
    parent = sc.parent.copy()
    parent[parent == bb.DETACHED] = scenes.NO_PARENT          # a parent the hierarchy does not know: falls back to true
    vis = rng.choice([orc.VIS_INHERITED] * 3 + [orc.VIS_HIDDEN, orc.VIS_VISIBLE], n).astype(np.uint8)
    vis[rng.random(n) < 0.03] |= orc.VIS_NO_COMPONENTS
    pipe = bb.VisibilityPipeline(sc)
    world = OracleWorld(sc, True)
    try:
        inh = (sc.flags & 1).astype(np.uint8)
        edited = np.arange(n, dtype=np.uint32)               # first run: every Visibility is Added
        for step in range(6):
            if step:
                edited = np.unique(rng.integers(0, n, 25)).astype(np.uint32)
                vis[edited] = (vis[edited] & 4) | rng.choice([0, 1, 2], len(edited)).astype(np.uint8)
            pipe.ctx.upload_visibility(0, vis)
            pipe.ctx.propagate_visibility()
            want, want_ch = orc.visibility_propagate(parent, vis, inh, edited)
            got, got_ch = pipe.ctx.download_inherited_visibility(0, n)
            assert (got == want).all(), f"step {step}: InheritedVisibility differs on rows {np.nonzero(got != want)[0][:8]}"
            assert (got_ch == want_ch).all(), f"step {step}: change flags differ on rows {np.nonzero(got_ch != want_ch)[0][:8]}"
            assert step == 0 or want_ch.any()
            inh = want
            # the cull phase sees the new column
            sc.flags = ((sc.flags & ~np.uint8(1)) | inh).astype(np.uint8)
            if step:
                scenes.advance_cameras(sc, 0.05)
            pipe.update_views()
            compare_frame(pipe, world, step)
    finally:
        pipe.close()


# ---- N3: check_point_light_mesh_visibility (shadow-view culling of point lights) --------------------------------
@pytest.mark.parametrize("seed,shuffle", [(8, True), (9, False)])
def test_n3_point_light_shadow_culling(seed, shuffle):
    sc = _random_scene(seed, n_roots=90, n_lights=20, shuffle_entities=shuffle)
    rng = np.random.default_rng(seed)
    n = sc.n
    sc.shadow_lights = np.sort(rng.choice(len(sc.light_row), 9, replace=False)).astype(np.uint32)   # shadow_maps_enabled
    sc.shadow_caster = (rng.random(n) < 0.8).astype(np.uint8)
    sc.shadow_caster[sc.light_row] = 0                       # lights are not Mesh3d
    sc.shadow_near_z = 0.1
    sc.shadow_lod_origin = 0                                 # view 0 is the shadow LOD origin
    pipe = bb.VisibilityPipeline(sc)
    world = OracleWorld(sc, True)
    try:
        pipe.ctx.upload_shadow_casters(0, sc.shadow_caster)
        seen = 0
        for f in range(5):
            if f:
                scenes.advance_cameras(sc, 0.15)
                rows = np.unique(rng.integers(0, n, n // 25)).astype(np.uint32)
                sc.trs[rows, 0:3] += rng.uniform(-2, 2, (len(rows), 3)).astype(np.float32)
                pipe.ctx.upload_transforms_scattered(rows, sc.trs[rows])
                world.tchanged[rows] = 1
                sc.view_flags = [bb.VIEW_ACTIVE, 0 if f == 3 else bb.VIEW_ACTIVE, bb.VIEW_ACTIVE]
            pipe.update_views()
            compare_frame(pipe, world, f)
            seen += sum(len(l) for six in world.shadow_result.values() for l in six)
            assert len(world.shadow_result) > 0
        assert seen > 100       # the lists are not trivially empty
    finally:
        pipe.close()


def test_n3_rows_only_lights_see_become_visible():
    """A mesh outside every camera frustum but inside a shadow light's range: ViewVisibility comes from set_visible()
    of the light pass alone, with the change flag firing on the hidden -> visible transition only."""
    sc = scenes.forest(n_trees=80, levels=5, n_lights=6)
    sc.trs[sc.roots, 0:3] *= np.float32(0.12)                # pull the trees inside the lights' reach
    sc.light_range[:] = 45.0
    sc.bounds[sc.light_row, 3] = 45.0
    sc.shadow_lights = np.arange(6, dtype=np.uint32)
    sc.shadow_caster = np.ones(sc.n, np.uint8); sc.shadow_caster[sc.light_row] = 0
    sc.shadow_near_z = 0.1
    sc.shadow_lod_origin = -1
    sc.cameras = sc.cameras[:1]                              # one camera: most meshes are outside its frustum
    pipe = bb.VisibilityPipeline(sc)
    world = OracleWorld(sc, True)
    try:
        pipe.ctx.upload_shadow_casters(0, sc.shadow_caster)
        for f in range(4):
            if f:
                scenes.advance_cameras(sc, 0.3)
            pipe.update_views()
            compare_frame(pipe, world, f)
        vv, _ = pipe.ctx.download_view_visibility(0, sc.n)
        cam = pipe.ctx.download_visible(0)
        assert (vv & 1).sum() > len(cam)                     # rows visible to lights only
    finally:
        pipe.close()


@pytest.mark.parametrize("seed,shuffle", [(12, True), (13, False)])
def test_n3_spot_lights_and_directional_cascades(seed, shuffle):
    """The other two halves of shadow-view culling on the device (b200vis_set_shadow_items): spot lights
    (bevy_light/src/lib.rs:670-749: one frustum, near + far planes, range-sphere pre-test, only lights some view lists) and
    directional-light cascades (lib.rs:342-510: one frustum per cascade, near plane skipped, gated on the VIEW's range bit),
    mixed with point lights in one pass.  Lists, ViewVisibility and its change flags against the oracle, frame after frame."""
    import oracle as orc
    sc = _random_scene(seed, n_roots=90, n_lights=20, shuffle_entities=shuffle)
    rng = np.random.default_rng(seed)
    n = sc.n
    caster = (rng.random(n) < 0.8).astype(np.uint8)
    caster[sc.light_row] = 0
    lights = rng.permutation(len(sc.light_row))
    spot_ords, point_ords = np.sort(lights[:6]), np.sort(lights[6:10])
    ident = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32)
    pipe = bb.VisibilityPipeline(sc)
    world = OracleWorld(sc, True)
    V = len(sc.cameras)
    try:
        pipe.ctx.upload_shadow_casters(0, caster)
        seen = 0
        for f in range(4):
            if f:
                scenes.advance_cameras(sc, 0.2)
                rows = np.unique(rng.integers(0, n, n // 20)).astype(np.uint32)
                sc.trs[rows, 0:3] += rng.uniform(-2, 2, (len(rows), 3)).astype(np.float32)
                pipe.ctx.upload_transforms_scattered(rows, sc.trs[rows])
                world.tchanged[rows] = 1
            pipe.update_views()
            planes = np.stack([np.ctypeslib.as_array(v.half_spaces).reshape(6, 4).copy() for v in pipe.views])
            # ---- oracle: check_visibility with mark_newly_hidden deferred, the light passes, then mark_newly_hidden
            rc, _ = orc.propagate(sc.parent, sc.trs, world.gt, world.tchanged, True)
            world.tchanged[:] = 0
            orc.set_defer_mark_newly_hidden(True)
            try:
                vv_changed, lists = orc.cull(world.gt, sc.bounds, sc.flags, sc.class_mask, sc.entity_bits, world.vv, planes,
                                             view_layers=sc.view_layers, view_flags=sc.view_flags, layer_mask=sc.layer_mask,
                                             range_mask=sc.range_mask, view_range_index=sc.view_range_index)
            finally:
                orc.set_defer_mark_newly_hidden(False)
            lists = [l if l is not None else world.last_lists[v] for v, l in enumerate(lists)]
            world.last_lists = lists
            listed = set(np.concatenate(lists).tolist())
            # ---- this frame's items: frusta from the lights' GlobalTransforms of this frame
            items, oracle_jobs = [], []
            for o in spot_ords:
                row = int(sc.light_row[o])
                fr = orc.point_light_frusta(world.gt[row], sc.light_range[o], 0.1)[int(o) % 6]    # any single frustum at the light
                ll = 1 if sc.light_layers is None else int(sc.light_layers[o])
                items.append(dict(kind=1, light_row=row, range=float(sc.light_range[o]), range_view_index=0, layer_mask=ll, frusta=fr))
                oracle_jobs.append(("spot", row, o, fr, ll))
            for o in point_ords:
                row = int(sc.light_row[o])
                fr = orc.point_light_frusta(world.gt[row], sc.light_range[o], 0.1)
                ll = 1 if sc.light_layers is None else int(sc.light_layers[o])
                items.append(dict(kind=0, light_row=row, range=float(sc.light_range[o]), range_view_index=0, layer_mask=ll, frusta=fr))
                oracle_jobs.append(("point", row, o, fr, ll))
            casc = []
            for v in range(min(V, 2)):                            # one directional light, cascades of views 0 and 1
                for c, rr in enumerate((25.0, 80.0)):
                    centre = np.asarray(sc.cameras[v].gt[9:12], np.float32) + np.float32(5.0 * c)
                    fr = orc.point_light_frusta(np.concatenate([ident, centre]).astype(np.float32), rr, 0.1)[(v + c) % 6]
                    vri = -1 if sc.view_range_index is None else int(sc.view_range_index[v])
                    items.append(dict(kind=2, range_view_index=vri, layer_mask=3, frusta=fr))
                    casc.append((v, fr, vri))
            pipe.ctx.run(bb.STAGE_ALL if len(sc.light_row) else (bb.STAGE_PROPAGATE | bb.STAGE_CULL))
            pipe.ctx.set_shadow_items(items)
            pipe.ctx.run_shadow_culling()
            # ---- oracle light passes in the reference's order: directional first (lib.rs:342), then point / spot (:517)
            want = {}
            dir_items = []
            for v in range(min(V, 2)):
                frs = np.stack([fr for (vv_, fr, _) in casc if vv_ == v])
                dir_items.append((frs, 3, casc[[i for i, c_ in enumerate(casc) if c_[0] == v][0]][2]))
            got_dir = orc.check_dir_light_mesh_visibility(world.gt, sc.bounds, sc.flags, caster, sc.entity_bits, world.vv, vv_changed,
                                                          dir_items, layer_mask=sc.layer_mask, range_mask=sc.range_mask)
            k = len(spot_ords) + len(point_ords)
            for lists_of_item in got_dir:
                for rows_ in lists_of_item:
                    want[(k, 0)] = rows_; k += 1
            for i, (kind, row, o, fr, ll) in enumerate(oracle_jobs):
                if row not in listed:                               # the light is in no view's VisibleEntities: not processed
                    for face in range(6):
                        want[(i, face)] = np.zeros(0, np.uint32)
                    continue
                sphere = np.concatenate([world.gt[row, 9:12], [sc.light_range[o]]]).astype(np.float32)[None]
                if kind == "spot":
                    r = orc.check_spot_light_mesh_visibility(world.gt, sc.bounds, sc.flags, caster, sc.entity_bits, world.vv, vv_changed,
                                                             sphere, fr[None], layer_mask=sc.layer_mask, range_mask=sc.range_mask,
                                                             lod_origin_index=0, light_layers=np.array([ll], np.uint64))
                    want[(i, 0)] = r[0]
                else:
                    r = orc.check_point_light_mesh_visibility(world.gt, sc.bounds, sc.flags, caster, sc.entity_bits, world.vv, vv_changed,
                                                              sphere, fr[None], layer_mask=sc.layer_mask, range_mask=sc.range_mask,
                                                              lod_origin_index=0, light_layers=np.array([ll], np.uint64))
                    for face in range(6):
                        want[(i, face)] = r[0][face]
            orc.mark_newly_hidden(sc.flags, world.vv, vv_changed)
            # ---- compare
            for (i, face), rows_ in want.items():
                got = pipe.ctx.download_shadow_visible(i, face)
                assert len(got) == len(rows_) and (got == rows_).all(), f"frame {f} item {i} face {face}: {len(got)} vs {len(rows_)}"
                seen += len(rows_)
            vv, vch = pipe.ctx.download_view_visibility(0, n)
            assert (vv == world.vv).all(), f"frame {f}: ViewVisibility differs on rows {np.nonzero(vv != world.vv)[0][:8]}"
            assert (vch == vv_changed).all(), f"frame {f}: Changed<ViewVisibility> differs"
            pipe.read_feedback()
        assert seen > 200
    finally:
        pipe.close()
