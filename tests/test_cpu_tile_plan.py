"""Host-side planner of the default (CTA-per-tile) kernel, checked on CPU.

The tile kernels (k_propagate_cull_lean, k_propagate_cull_tma) hand a tile's levels over through hardware NAMED BARRIERS: barrier l is joined by exactly the warps
that hold a row of level l-1 (producers: bar.arrive, or bar.sync when they also consume) or of level l (consumers:
bar.sync), with the participant count the planner wrote into Tile::lvl_warps.  A count that disagrees with what the warps
derive from their own rows' topo words is a hang on the device, so the protocol is replayed here: every warp runs the
kernel's loop over the plan (same masks, same order) against a model of the barriers, and must come out the other end with
every parent computed before its children."""
import numpy as np
import pytest

import bevy_b200 as bb
from bevy_b200 import abi, scenes

NO_PARENT, DETACHED = 0xFFFFFFFF, 0xFFFFFFFE
T_DETACHED = 1 << 31


def replay_tile(base, nr, n_levels, wsm, lvl_warps, topo, top_k=0):
    """Warps as coroutines over the kernel's level loop; returns the order in which rows were computed.
    top_k >= 2 (kernel 1L): the rows of in-tile depth < top_k are walked in registers by warp 0 before the loop -- a child takes
    its parent's matrix with a warp shuffle from lane `parent`, so all of them must sit among the tile's first 32 rows -- and
    levels 1 .. top_k - 1 drop out of every warp's schedule."""
    n_warps = 8
    local = topo[base:base + nr]
    depth = ((local >> 9) & 0x1FF).astype(np.int64)
    level_of = depth                 # a detached row has depth 0: it publishes "never visited" before the loop
    lmask = [0] * n_warps
    for i in range(nr):
        assert level_of[i] < 16
        lmask[i >> 5] |= 1 << int(level_of[i])
    done = np.zeros(nr, bool)
    done[level_of == 0] = True       # level 0 is computed before the loop
    if top_k >= 2:
        for d in range(1, top_k):    # the shuffle walk: level by level inside warp 0
            for i in np.nonzero(level_of == d)[0]:
                p = int(local[i] & 0x1FF)
                assert i < 32 and p < 32, f"row {base + i} of top level {d} (parent {base + p}) is outside the first warp"
                if not (int(local[i]) & T_DETACHED):
                    assert done[p] and level_of[p] == d - 1
                done[i] = True
    # per warp: the list of (level, role) steps the kernel's while loop takes
    steps = []
    for w in range(n_warps):
        need = (lmask[w] | (lmask[w] << 1)) & ((1 << n_levels) - 2)
        if top_k >= 2:
            need &= ~((1 << top_k) - 2)
        st = []
        for lvl in range(1, n_levels):
            if not (need >> lvl) & 1:
                continue
            consumer = bool((lmask[w] >> lvl) & 1)
            if (wsm >> lvl) & 1:
                if consumer:
                    st.append((lvl, "warp"))
            else:
                st.append((lvl, "sync" if consumer else "arrive"))
        steps.append(st)
    pc = [0] * n_warps
    arrived = {}                     # barrier id -> set of warps that have arrived in the current generation
    waiting = {}                     # warp -> barrier id it is blocked on
    order = []

    def compute(w, lvl):
        for i in range(32 * w, min(32 * w + 32, nr)):
            if level_of[i] == lvl:
                p = int(local[i] & 0x1FF)
                assert done[p], f"row {base + i} (level {lvl}) computed before its parent {base + p}"
                done[i] = True
                order.append(i)

    progress = True
    while progress:
        progress = False
        for w in range(n_warps):
            if w in waiting or pc[w] >= len(steps[w]):
                continue
            lvl, role = steps[w][pc[w]]
            if role == "warp":
                compute(w, lvl); pc[w] += 1; progress = True
                continue
            cnt = (lvl_warps >> (4 * lvl)) & 15
            assert cnt >= 1, f"level {lvl}: a warp joins a barrier the planner gave no participants"
            arrived.setdefault(lvl, set()).add(w)
            progress = True
            if role == "arrive":
                pc[w] += 1
            else:
                waiting[w] = lvl
            if len(arrived[lvl]) == cnt:      # generation complete: release the consumers
                for x in [x for x, b in waiting.items() if b == lvl]:
                    del waiting[x]
                    compute(x, lvl); pc[x] += 1
                arrived[lvl] = set()
            assert len(arrived[lvl]) < max(cnt, 1) or cnt == 0
    assert not waiting, f"deadlock: warps {sorted(waiting)} wait on barriers {sorted(set(waiting.values()))}"
    assert all(pc[w] == len(steps[w]) for w in range(n_warps))
    assert all(not s for s in arrived.values()), "a barrier generation was left half full (the next tile would inherit it)"
    assert done.all()


def check(parent, tile_rows=0):
    parent = np.asarray(parent, np.uint32)
    desc, topo = abi.host_tile_plan(parent, tile_rows)
    named = 0
    for base, nr, n_levels, wsm, top, lo, hi, _pass in desc.tolist():
        lvl_warps = lo | (hi << 32)
        if not (2 <= n_levels <= 8):
            assert lvl_warps == 0
            continue
        if lvl_warps == 0:
            continue
        named += 1
        # the count the kernel will use = warps holding a row of level l-1 or l
        local = topo[base:base + nr]
        lv = (local >> 9) & 0x1FF
        for l in range(1, n_levels):
            warps = {i >> 5 for i in range(nr) if lv[i] in (l - 1, l)}
            assert (lvl_warps >> (4 * l)) & 15 == len(warps), f"tile at {base}: level {l}"
        assert lvl_warps >> (4 * n_levels) == 0 and lvl_warps & 15 == 0
        replay_tile(base, nr, n_levels, wsm, lvl_warps, topo)
        if top >= 2:                 # the default kernel's schedule: top levels in registers, the rest through the barriers
            replay_tile(base, nr, n_levels, wsm, lvl_warps, topo, top_k=top)
    return desc, named


def test_bench_forest_levels_meet_at_named_barriers():
    sc = scenes.forest(n_trees=20, levels=8, n_lights=5)
    desc, named = check(sc.parent)
    trees = desc[desc[:, 1] == 255]
    assert named >= 19 and (trees[:, 2] == 8).all()
    # BFS tree of 255 (level l = rows 2^l - 1 .. 2^(l+1) - 2): levels 0..4 live in warp 0; level 5 = row 31 (warp 0) + warp 1;
    # level 6 = warps 1..3; level 7 = warps 3..7
    lw = int(trees[0, 5]) | (int(trees[0, 6]) << 32)
    assert [(lw >> (4 * l)) & 15 for l in range(1, 8)] == [1, 1, 1, 1, 2, 4, 7]
    assert (trees[:, 4] == 5).all()           # top_levels: rows 0..30 = levels 0..4 are walked in registers by warp 0


def test_flat_deep_and_wide_shapes():
    check(np.full(1000, NO_PARENT, np.uint32))
    chain = np.concatenate([[NO_PARENT], np.arange(699)]).astype(np.uint32)
    desc, named = check(chain)
    assert named == 0                         # 129-level tiles keep the CTA-wide walk
    fan = np.concatenate([[NO_PARENT], np.zeros(254, np.uint32)]).astype(np.uint32)
    check(fan)
    check(scenes.propagate_bench_scene().parent)
    check(scenes.many_cubes(5000, n_lights=16).parent)


@pytest.mark.parametrize("seed", range(8))
def test_random_forests(seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(50, 4000))
    parent = np.full(n, NO_PARENT, np.uint32)
    for r in range(1, n):
        k = rng.random()
        if k < 0.12:
            continue
        if k < 0.15:
            parent[r] = DETACHED
            continue
        lo = max(0, r - int(rng.integers(1, 300)))
        parent[r] = rng.integers(lo, r)
    order = bb.plan_row_order(parent)
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    p2 = parent[order].copy()
    m = p2 < n
    p2[m] = inv[p2[m]].astype(np.uint32)
    check(p2)
    check(p2, 64)
    check(parent)                             # un-renumbered (topological but interleaved) rows too
