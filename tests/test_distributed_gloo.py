"""world_size=2 over gloo on CPU: the host-side multi-GPU logic (row-range sharding by whole trees, the
rank-major slab all-gather, the feedback reduction, global light ordinals).  Each rank plays its GPU with
the CPU oracle; the gathered result must equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
from bevy_b200 import parallel, scenes

WORDS, MAXC = 2, 4096          # 64 lights per rank capacity


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _views(sc):
    return np.stack([orc.compute_frustum(orc.perspective(c.fov, c.aspect, c.near), c.gt, c.far) for c in sc.cameras])


def _rank_frame(sc):
    """propagate + cull + per-view cluster assignment of this shard's own lights (CPU oracle)."""
    n = sc.n
    gt = np.tile(orc.IDENTITY_GT, (n, 1)); vv = np.zeros(n, np.uint8)
    orc.propagate(sc.parent, sc.trs, gt, np.ones(n, np.uint8))
    planes = _views(sc)
    _, lists = orc.cull(gt, sc.bounds, sc.flags, sc.class_mask, sc.entity_bits, vv, planes)
    vis = np.nonzero(vv[sc.light_row] & 1)[0]
    lights = np.concatenate([gt[sc.light_row[vis], 9:12], sc.light_range[vis, None]], 1).astype(np.float32)
    per_view = []
    for v, cam in enumerate(sc.cameras):
        cfv = orc.perspective(cam.fov, cam.aspect, cam.near)
        out, off, idx, _ = orc.assign_lights_to_clusters(orc.default_cluster_view_in(cam.gt, cfv, planes[v]), lights)
        per_view.append((out, off, vis[idx]))
    return lists, per_view


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = scenes.forest(n_trees=9, levels=5, n_lights=21)
    sub, rows, (l_lo, l_hi) = parallel.shard_scene(full, rank, world, per_tree=31)
    lists, per_view = _rank_frame(sub)
    V = len(full.cameras)
    # this rank's slab: [V][WORDS][MAXC] bit = local light ordinal
    slab = np.zeros((V, WORDS, MAXC), np.uint32)
    for v, (out, off, idx) in enumerate(per_view):
        nc = out.dims[0] * out.dims[1] * out.dims[2]
        for c in range(nc):
            for l in idx[off[c]:off[c + 1]]:
                slab[v, l >> 5, c] |= np.uint32(1 << (l & 31))
    send = torch.from_numpy(slab.view(np.int32).reshape(-1))
    recv = torch.zeros(world * send.numel(), dtype=torch.int32)
    parallel.all_gather_slabs(recv, send)                       # the single data-path collective
    far, cnt = parallel.reduce_feedback([p[0].farthest_z for p in per_view], [p[0].total_index_count for p in per_view])
    if rank == 0:
        ret["recv"] = recv.numpy().view(np.uint32).reshape(world, V, WORDS, MAXC).copy()
        ret["far"], ret["cnt"] = far, cnt
        ret["dims"] = [tuple(p[0].dims) for p in per_view]
    # visible lists: per-rank lists are sorted and disjoint; the merge is the global list
    gathered = [None] * world
    dist.all_gather_object(gathered, [rows[l].tolist() for l in lists])
    if rank == 0:
        ret["visible"] = [parallel.merge_visible_lists([g[v] for g in gathered]).tolist() for v in range(V)]
    dist.destroy_process_group()


def test_two_rank_shard_gather_equals_single_process():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    full = scenes.forest(n_trees=9, levels=5, n_lights=21)
    lists, per_view = _rank_frame(full)
    light_ranges = parallel.shard_bounds(21, world)
    assert light_ranges == [(0, 11), (11, 21)]
    for v, (out, off, idx) in enumerate(per_view):
        assert ret["visible"][v] == lists[v].tolist()
        assert tuple(out.dims) == ret["dims"][v]
        recv = ret["recv"]
        nc = out.dims[0] * out.dims[1] * out.dims[2]
        for c in range(nc):
            got = []
            for r in range(world):                                  # rank-major, then ascending local ordinal
                for w in range(WORDS):
                    m = int(recv[r, v, w, c])
                    got += [r * WORDS * 32 + w * 32 + b for b in range(32) if (m >> b) & 1]
            glob = parallel.global_light_ordinal(np.array(got, np.int64), WORDS * 32, light_ranges) if got else np.zeros(0, np.int64)
            assert glob.tolist() == idx[off[c]:off[c + 1]].tolist(), (v, c)
        assert abs(ret["far"][v] - out.farthest_z) < 1e-6 * max(1.0, abs(out.farthest_z))
        assert ret["cnt"][v] == out.total_index_count
