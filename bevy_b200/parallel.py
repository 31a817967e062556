"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

The path shards by contiguous row range (whole trees per rank), so propagate and cull need no
collective.  The only data-path exchange is ONE all-gather per frame of the fixed-size
cluster x light bitmask slabs; the two per-view feedback scalars
(Clusters::last_frame_*) are reduced with the frame statistics.  Works with the NCCL backend on
GPUs and with gloo on CPU tensors (the tests run world_size=2 over gloo).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import scenes


def shard_bounds(n_units, world):
    """[lo, hi) unit ranges per rank: contiguous, sizes differ by at most one (like
    QueryState::par_fold_init_unchecked_manual's batching, crates/bevy_ecs/src/query/state.rs:1661-1676)."""
    base, rem = divmod(n_units, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_scene(scene, rank, world, per_tree):
    """This rank's contiguous row range of a forest scene whose rows are [trees | lights]: whole trees
    (a 255-node tree never spans GPUs) plus a contiguous slice of the lights."""
    n_lights = len(scene.light_row)
    n_tree_rows = scene.n - n_lights
    assert n_tree_rows % per_tree == 0
    t_lo, t_hi = shard_bounds(n_tree_rows // per_tree, world)[rank]
    l_lo, l_hi = shard_bounds(n_lights, world)[rank]
    rows = np.concatenate([np.arange(t_lo * per_tree, t_hi * per_tree), n_tree_rows + np.arange(l_lo, l_hi)])
    remap = np.full(scene.n, -1, np.int64)
    remap[rows] = np.arange(len(rows))
    parent = scene.parent[rows].astype(np.int64)
    has_parent = parent < scene.n
    parent[has_parent] = remap[parent[has_parent]]
    assert (parent[has_parent] >= 0).all(), "a tree spans two shards"
    roots = None
    if scene.roots is not None:
        sel = (remap[scene.roots] >= 0)
        roots = remap[scene.roots[sel]].astype(np.uint32)
    sub = scenes.Scene(f"{scene.name}[{rank}/{world}]", parent.astype(np.uint32), scene.trs[rows].copy(),
                       scene.bounds[rows].copy(), scene.flags[rows].copy(), scene.class_mask[rows].copy(),
                       scene.entity_bits[rows].copy(), remap[scene.light_row[l_lo:l_hi]].astype(np.uint32),
                       scene.light_range[l_lo:l_hi].copy(), scene.cameras, roots, scene.screen)
    return sub, rows, (l_lo, l_hi)


def all_gather_slabs(recv, send, group=None):
    """The single collective of the data path: rank-major concatenation of the per-rank slabs."""
    dist.all_gather_into_tensor(recv, send, group=group)


def reduce_feedback(farthest_z, index_count, group=None, device=None):
    """Clusters::last_frame_farthest_z / last_frame_total_cluster_index_count must be identical on every
    rank: max of the per-rank maxima, sum of the per-rank counts (both per view)."""
    far = torch.as_tensor(np.asarray(farthest_z, np.float32), device=device)
    cnt = torch.as_tensor(np.asarray(index_count, np.int64), device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(far, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return far.cpu().numpy(), cnt.cpu().numpy()


def global_light_ordinal(local_ordinal, max_lights_per_rank, light_ranges):
    """Cluster lists carry rank-major ordinals r * max_lights_per_rank + local; map them to the global
    light index given each rank's [lo, hi) light slice."""
    local_ordinal = np.asarray(local_ordinal, np.int64)
    r = local_ordinal // max_lights_per_rank
    l = local_ordinal % max_lights_per_rank
    lo = np.array([a for a, _ in light_ranges], np.int64)
    return lo[r] + l


def merge_visible_lists(entity_bits_per_rank):
    """Global VisibleEntities of one view from the per-rank lists.  Each rank's list is already ascending in
    Entity::to_bits(); ranks own disjoint entity sets, so a k-way merge (here: sort of the concatenation)
    gives the list the reference's serial sort_unstable produces (visibility/mod.rs:870-874)."""
    cat = np.concatenate([np.asarray(x, np.uint64) for x in entity_bits_per_rank]) if entity_bits_per_rank else np.zeros(0, np.uint64)
    return np.sort(cat, kind="stable")
