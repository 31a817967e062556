"""CPU-side checks (no GPU): the C-ABI library loads and exports every declared symbol, the host-side
per-view maths of the product equals the oracle's, the z-slice thresholds reproduce libm's view_z_to_z_slice,
the row-order planner, the scene generators, and the ABI fails loudly without a CUDA device."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import bevy_b200 as bb
from bevy_b200 import abi, scenes
import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200vis.h")).read()
    declared = set(re.findall(r"B200VIS_API\s+[\w\s\*]+?\b(b200vis_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = bb.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/b200vis.h but not exported"
    assert declared == set(abi.EXPORTED_SYMBOLS), declared ^ set(abi.EXPORTED_SYMBOLS)
    assert bb.abi_version() == 2


def test_create_without_cuda_fails_loudly():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(bb.B200VisError) as e:
        bb.Context(16)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_struct_sizes_match_the_header():
    sizes = (C.c_uint32 * 6)()
    bb.load_library().b200vis_struct_sizes(sizes)
    py = [abi.Config, abi.View, abi.ClusterView, abi.FrameStats, abi.ClusterConfig, abi.ClusterFeedback]
    assert list(sizes) == [C.sizeof(s) for s in py]


@pytest.mark.parametrize("yaw,pitch,pos,scale", [(0.0, 0.0, (0, 0, 0), 1.0), (1.1, -0.3, (5, -2, 9), 1.0),
                                                 (2.9, 0.7, (-40, 3, 2), 2.5), (4.0, 0.1, (0.5, 0.5, 0.5), 0.25)])
def test_host_frustum_and_cluster_setup_equal_the_oracle(yaw, pitch, pos, scale):
    q = scenes.quat_mul(scenes.quat_axis("y", yaw), scenes.quat_axis("x", pitch))
    gt = scenes.quat_to_gt(q, pos)
    gt[:9] *= np.float32(scale)
    for fov, aspect, near, far in [(math.pi / 4, 16 / 9, 0.1, 1000.0), (1.2, 1.0, 0.5, 50.0)]:
        cfv_p = bb.host_perspective(fov, aspect, near)
        cfv_o = orc.perspective(fov, aspect, near)
        assert (cfv_p.view(np.uint32) == cfv_o.view(np.uint32)).all()
        fr_p = bb.host_compute_frustum(cfv_p, gt, far)
        fr_o = orc.compute_frustum(cfv_o, gt, far)
        assert (fr_p.view(np.uint32) == fr_o.view(np.uint32)).all()
        for fb_far, fb_cnt in [(None, None), (37.5, 100), (12.25, 50000)]:
            fb = abi.ClusterFeedback()
            if fb_far is not None:
                fb.has_farthest_z, fb.farthest_z, fb.has_index_count, fb.index_count = 1, fb_far, 1, fb_cnt
            cv, scratch = bb.host_cluster_view_setup(bb.host_default_cluster_config(1920, 1080), gt, cfv_p, fr_p, 1, fb)
            vin = orc.default_cluster_view_in(gt, cfv_o, fr_o, last_farthest_z=fb_far, last_index_count=fb_cnt)
            out, _, _, planes = orc.assign_lights_to_clusters(vin, np.zeros((0, 4), np.float32), want_planes=True)
            assert tuple(cv.dims) == tuple(out.dims) and tuple(cv.tile_size) == tuple(out.tile_size)
            for a, b in [(cv.near_z, out.near), (cv.far_z, out.far), (cv.cluster_factors[0], out.cluster_factors[0]),
                         (cv.cluster_factors[1], out.cluster_factors[1]), (cv.view_from_world_scale_max, out.view_from_world_scale_max)]:
                assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
            assert (np.array(cv.view_from_world[:], np.float32).view(np.uint32) == np.array(out.view_from_world[:], np.float32).view(np.uint32)).all()
            for tab, ref, cnt in [(cv.x_planes, planes[0], cv.dims[0] + 1), (cv.y_planes, planes[1], cv.dims[1] + 1),
                                  (cv.z_planes, planes[2], cv.dims[2] + 1)]:
                got = np.ctypeslib.as_array(tab, shape=(cnt * 4,)).view(np.uint32)
                assert (got == ref.reshape(-1).view(np.uint32)).all()


def test_cluster_config_none_and_empty_viewport_clear():
    cfg = bb.host_default_cluster_config(0, 1080)
    cv, _ = bb.host_cluster_view_setup(cfg, orc.IDENTITY_GT, bb.host_perspective(1, 1, 0.1), np.zeros((6, 4), np.float32))
    assert cv.enabled == 0
    cfg = bb.host_default_cluster_config(800, 600); cfg.kind = 0
    cv, _ = bb.host_cluster_view_setup(cfg, orc.IDENTITY_GT, bb.host_perspective(1, 1, 0.1), np.zeros((6, 4), np.float32))
    assert cv.enabled == 0


def test_plan_row_order_is_topological_and_tree_contiguous():
    rng = np.random.default_rng(3)
    n = 2000
    parent = np.full(n, bb.NO_PARENT, np.uint32)
    for r in range(1, n):                     # arbitrary forest: parents may come AFTER children in row order
        if rng.random() < 0.9:
            parent[r] = rng.integers(0, r)
    perm = rng.permutation(n)                 # scramble rows
    inv = np.argsort(perm)
    p2 = np.full(n, bb.NO_PARENT, np.uint32)
    has = parent != bb.NO_PARENT
    p2[perm[has]] = perm[parent[has]]
    order = bb.plan_row_order(p2)             # new -> old
    assert sorted(order.tolist()) == list(range(n))
    new_of_old = np.argsort(order)
    for new, old in enumerate(order):
        if p2[old] != bb.NO_PARENT:
            assert new_of_old[p2[old]] < new  # parent row precedes the child
    # every tree is one contiguous block
    root_of = np.arange(n)
    for new, old in enumerate(order):
        if p2[old] != bb.NO_PARENT:
            root_of[old] = root_of[p2[old]]
    blocks = root_of[order]
    changes = (blocks[1:] != blocks[:-1]).sum() + 1
    assert changes == len(np.unique(root_of))
    assert bb.plan_row_order(np.array([1, 0], np.uint32)) is not None if False else True
    with pytest.raises(bb.B200VisError):
        bb.plan_row_order(np.array([1, 0], np.uint32))       # a cycle: rows never reached


def test_scene_generators_match_the_config_sizes():
    sc = scenes.propagate_bench_scene()
    assert sc.n == 48 * 1077 + 12000                       # benches/.../propagate.rs:23-24,74
    sc = scenes.forest(n_trees=10, levels=8, n_lights=5)
    assert sc.n == 10 * 255 + 5 and (sc.parent[sc.light_row] == bb.NO_PARENT).all()
    has = sc.parent != bb.NO_PARENT
    assert (sc.parent[has] < np.nonzero(has)[0]).all()     # BFS per tree => topological
    sc = scenes.many_cubes(1000)
    assert np.allclose(np.linalg.norm(sc.trs[:, 0:3], axis=1), 500.0, rtol=1e-5)
    assert np.allclose(np.linalg.norm(sc.trs[:, 3:7], axis=1), 1.0, atol=1e-6)


def test_z_slice_thresholds_reproduce_libm_view_z_to_z_slice():
    """The device never calls logf: it counts host-computed thresholds.  That count must equal the reference's
    view_z_to_z_slice (oracle = the platform libm, as Rust's f32::ln) for every view_z, including the neighbours of
    every threshold, zero, negatives behind the camera, infinities and NaN."""
    import ctypes as C
    lib = orc.lib()
    rng = np.random.default_rng(9)
    cases = []
    for _ in range(40):
        near = float(rng.uniform(0.05, 20)); far = near * float(np.exp(rng.uniform(0.0, 7.0)))
        zs = int(rng.integers(1, 40))
        k = (np.float32(zs) - np.float32(1)) / np.float32(orc.lib().orc_logf(np.float32(far) / np.float32(near)))
        cases.append((np.array([k, np.float32(orc.lib().orc_logf(np.float32(near))) * k], np.float32), zs, False))
        cases.append((np.array([-near, zs / (-far + near)], np.float32), zs, True))
    cases.append((np.array([np.inf, np.nan], np.float32), 24, False))       # far == near degenerate
    cases.append((np.array([0.0, 0.0], np.float32), 1, False))
    for factors, zs, ortho in cases:
        thr = bb.host_z_slice_thresholds(factors, zs, ortho)
        fin = thr[np.isfinite(thr)]
        probe = [0.0, -0.0, 1e-30, 1e30, np.inf, -np.inf, np.nan, -1.0, -1e-3]
        for t_ in fin:
            probe += [t_, np.nextafter(np.float32(t_), np.float32(-np.inf)), np.nextafter(np.float32(t_), np.float32(np.inf))]
        probe += list(np.exp(rng.uniform(-5, 9, 200)))
        fp = factors.ctypes.data_as(C.POINTER(C.c_float))
        for u in np.array(probe, np.float32):
            want = lib.orc_view_z_to_z_slice(fp, zs, C.c_float(-u), int(ortho))
            got = 0 if np.isnan(u) else int(np.sum(u >= thr[~np.isnan(thr)]))
            got = min(got, zs - 1)
            assert got == want, (factors, zs, ortho, u, got, want)


def test_execution_plan_shapes():
    """The planner (no GPU needed): tiles, passes and error codes for the hierarchies the configs use."""
    sc = scenes.forest(n_trees=100, levels=8, n_lights=10)           # config #3 shape: one 255-node tree per tile
    tiles, passes, levels, ext = bb.host_plan_summary(sc.parent)
    assert (tiles, passes, levels, ext) == (101, 1, 8, 0)            # 100 trees + one tile of 10 flat light rows
    sc = scenes.propagate_bench_scene()                              # config #1: 1077-node trees span tiles
    tiles, passes, levels, ext = bb.host_plan_summary(sc.parent)
    assert passes == 3 and ext > 0 and tiles >= sc.n // 256
    flat = np.full(1000, bb.NO_PARENT, np.uint32)                    # config #2 shape: flat
    assert bb.host_plan_summary(flat) == (4, 1, 1, 0)
    chain = np.array([bb.NO_PARENT] + list(range(599)), np.uint32)   # a 600-deep chain: one pass per tile; a tile holds
    tiles, passes, levels, ext = bb.host_plan_summary(chain)         # at most 128 rows WITH in-tile children (+ the last child)
    assert tiles == 5 and passes == 5 and levels == 129 and ext == 4
    with pytest.raises(bb.B200VisError) as e:
        bb.host_plan_summary(np.array([1, 2, 0], np.uint32))
    assert e.value.code == 4                                         # cycle
    with pytest.raises(bb.B200VisError) as e:
        bb.host_plan_summary(np.array([bb.NO_PARENT, 7], np.uint32))
    assert e.value.code == 5                                         # parent out of range
    with pytest.raises(bb.B200VisError) as e:
        bb.host_plan_summary(np.array([1, bb.NO_PARENT], np.uint32))
    assert e.value.code == 8                                         # not in topological order
    assert bb.host_plan_summary(np.zeros(0, np.uint32)) == (0, 0, 0, 0)


def test_deep_tile_split_experiment_plan_structure():
    """B200VIS_SPLIT_DEEP_TILES=1 (off by default): a BFS-ordered 255-node tree is cut after its top five levels (31 rows);
    the bottoms become 3-level tiles with 32 external parents each, one pass later.  The flag is read once per process."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from bevy_b200 import abi, scenes; "
            "print(abi.host_plan_summary(scenes.forest(n_trees=10, levels=8, n_lights=0).parent)); "
            "print(abi.host_plan_summary(scenes.forest(n_trees=10, levels=4, n_lights=0).parent))") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for flag in ("0", "1"):
        env = dict(os.environ, B200VIS_SPLIT_DEEP_TILES=flag)
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        assert res.returncode == 0, res.stderr[-2000:]
        out[flag] = res.stdout.strip().splitlines()
    assert out["0"][0] == "(10, 1, 8, 0)"            # tiles, passes, max levels per tile, rows with an external parent
    assert out["1"][0] == "(20, 2, 5, 320)"
    assert out["0"][1] == out["1"][1]                # shallow trees are left alone
