"""Host-side planner of the default (warp-per-tile) kernel, checked on CPU: the schedule must be a valid execution order.

The kernel (bevy_b200/csrc/kernels.cu, k_tile_warp) walks a tile in chunks of 32 schedule slots and, inside a chunk, level
by level; a row reads its parent's GlobalTransform from the parent's shared-memory slot.  Invariants checked here for many
tree shapes: every row is scheduled exactly once; a row's in-tile parent sits in an earlier chunk, or in the same chunk at
a lower depth; a tile has at most 128 slot owners and slots are unique inside a tile; rows whose parent is in another tile
are in a later pass; the contiguity / non-root bit sets describe the schedule."""
import numpy as np
import pytest

import bevy_b200 as bb
from bevy_b200 import abi, scenes

NO_PARENT, DETACHED = 0xFFFFFFFF, 0xFFFFFFFE
T_ROOT, T_HAS_CHILDREN, T_EXT_PARENT, T_DETACHED, W_HAS_SLOT = 1 << 28, 1 << 29, 1 << 30, 1 << 31, 1 << 22


def check_plan(parent, tile_rows=0):
    parent = np.asarray(parent, np.uint32)
    n = len(parent)
    desc, nonroot, sched, wtopo = abi.host_warp_plan(parent, tile_rows)
    seen = np.zeros(n, np.int32)
    tile_of = np.full(n, -1, np.int64)
    pos_of = np.full(n, -1, np.int64)
    for ti, (base, nr, cc, _pass) in enumerate(desc):
        n_chunks, contig = int(cc) & 0xFF, int(cc) >> 8
        assert 1 <= nr <= 256 and n_chunks * 32 >= nr and n_chunks <= 8
        pad = 0x100 if nr == 256 else 0xFF
        slots = sched[ti, :n_chunks * 32].astype(np.int64)
        occupied = slots != pad
        rows = base + slots[occupied]
        assert len(rows) == nr and (np.sort(rows) == np.arange(base, base + nr)).all(), f"tile {ti}: schedule is not a permutation of its rows"
        assert (sched[ti, n_chunks * 32:] == 0xFF).all()
        seen[rows] += 1
        tile_of[rows] = ti
        pos_of[rows] = np.nonzero(occupied)[0]
        for c in range(n_chunks):
            lanes = np.nonzero(occupied[c * 32:(c + 1) * 32])[0]
            r = base + slots[c * 32 + lanes]
            is_contig = len(set((r - lanes).tolist())) <= 1
            assert bool((contig >> c) & 1) == is_contig, f"tile {ti} chunk {c}: contiguity bit"
            want = 0
            for lane, row in zip(lanes, r):
                if wtopo[row] & 0xFF:
                    want |= 1 << int(lane)
            assert int(nonroot[ti, c]) == want, f"tile {ti} chunk {c}: non-root bits"
    assert (seen == 1).all()
    depth = wtopo & 0xFF
    own = (wtopo >> 8) & 127
    pp = (wtopo >> 15) & 127
    has_kids = np.zeros(n, bool)
    for r in range(n):
        p = int(parent[r])
        w = int(wtopo[r])
        if p == NO_PARENT:
            assert w & T_ROOT and depth[r] == 0
        elif p == DETACHED:
            assert w & T_DETACHED and depth[r] == 0
        elif tile_of[p] == tile_of[r]:
            has_kids[p] = True
            assert depth[r] == depth[p] + 1
            assert wtopo[p] & W_HAS_SLOT and pp[r] == own[p], f"row {r}: parent slot"
            cr, cp = pos_of[r] // 32, pos_of[p] // 32
            assert cp < cr or (cp == cr and depth[p] < depth[r]), f"row {r}: parent scheduled too late"
        else:
            assert w & T_EXT_PARENT and depth[r] == 0
            assert desc[tile_of[p], 3] < desc[tile_of[r], 3], f"row {r}: parent's tile is not in an earlier pass"
            has_kids[p] = True
    assert (((wtopo & T_HAS_CHILDREN) != 0) == has_kids).all()
    for ti in range(len(desc)):
        base, nr = int(desc[ti, 0]), int(desc[ti, 1])
        w = wtopo[base:base + nr]
        owners = own[base:base + nr][(w & W_HAS_SLOT) != 0]
        assert len(owners) <= 128 and len(set(owners.tolist())) == len(owners), f"tile {ti}: slots"
    return desc


def test_binary_trees_get_one_chunk_aligned_tile_each():
    sc = scenes.forest(n_trees=20, levels=8, n_lights=5)
    desc = check_plan(sc.parent)
    trees = desc[desc[:, 1] == 255]
    assert len(trees) == 19          # the last tree shares its (full, unpadded) tile with the first light row
    assert ((trees[:, 2] & 0xFF) == 8).all() and ((trees[:, 2] >> 8) == 0xFF).all()      # 8 chunks, all contiguous
    _, _, sched, _ = abi.host_warp_plan(sc.parent)
    assert sched[0, 31] == 0xFF and sched[0, 32] == 31 and sched[0, 255] == 254            # level 5 starts chunk 1


def test_flat_rows_and_small_tiles():
    flat = np.full(1000, NO_PARENT, np.uint32)
    desc = check_plan(flat)
    assert len(desc) == 4 and (desc[:3, 1] == 256).all()
    desc = check_plan(flat, 32)
    assert len(desc) == 32 and (desc[:31, 1] == 32).all()


def test_config1_wide_trees_multi_pass():
    check_plan(scenes.propagate_bench_scene().parent)


def test_chain_is_cut_at_128_slot_owners():
    chain = np.concatenate([[NO_PARENT], np.arange(699)]).astype(np.uint32)
    desc = check_plan(chain)
    assert (desc[:, 1] <= 129).all()


@pytest.mark.parametrize("seed", range(6))
def test_random_forests(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(50, 3000))
    parent = np.full(n, NO_PARENT, np.uint32)
    for r in range(1, n):
        k = rng.random()
        if k < 0.15:
            continue                      # a new root / flat entity
        if k < 0.18:
            parent[r] = DETACHED
            continue
        lo = max(0, r - int(rng.integers(1, 400)))
        parent[r] = rng.integers(lo, r)  # topological by construction
    order = bb.plan_row_order(parent)    # BFS per tree, the layout the shim would upload
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    p2 = parent[order].astype(np.int64)
    real = p2 < n
    p2[real] = inv[p2[real]]
    check_plan(p2.astype(np.uint32))
    check_plan(p2.astype(np.uint32), 64)
    check_plan(parent)                   # and the unsorted (but topological) original
