"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/multi_gpu_parity.py
under torchrun with 2 ranks over NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["host_collective", "builtin", "builtin_slabs", "p2p", "p2p_slabs"])
def test_two_gpu_shard_gather_matches_oracle(mode):
    # builtin: the library's own ncclAllGather of the light-record blocks + the one-launch cluster stage on every rank;
    # builtin_slabs: the same collective over the cluster x light bit slabs (B200VIS_EXCHANGE_WHAT=slabs, the path that
    # remains for light counts beyond the fused kernel's shared memory); p2p / p2p_slabs: the same two payloads as peer stores
    # over NVLink (CUDA IPC) + per-frame stamps; host_collective: slabs by the host's own all-gather
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", {"builtin": "29533", "host_collective": "29534", "p2p": "29535", "builtin_slabs": "29536", "p2p_slabs": "29537"}[mode],
           os.path.join(here, "multi_gpu_parity.py")]
    if mode != "host_collective":
        cmd.append("--" + mode.split("_")[0])
    env = dict(os.environ)
    env.pop("B200VIS_EXCHANGE_WHAT", None)
    if mode.endswith("_slabs"):
        env["B200VIS_EXCHANGE_WHAT"] = "slabs"
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(res.stdout[-2000:]); print(res.stderr[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


def test_one_process_drives_two_gpus_through_peer_links():
    """A Bevy App is ONE process: two contexts (two devices) in this process, linked with b200vis_p2p_link (plain peer
    access, no IPC, no NCCL); one host thread calls run(ALL) on each per frame.  Merged visible lists and the cluster lists
    (built on BOTH devices from both devices' lights) against the single-process oracle on the whole scene."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    import bevy_b200 as bb
    from bevy_b200 import abi, parallel, scenes
    import oracle as orc
    from parity import OracleWorld
    world = 2
    per_tree, n_trees, n_lights = 63, 41, 37
    full = scenes.forest(n_trees=n_trees, levels=6, n_lights=n_lights)
    max_lights = max(hi - lo for lo, hi in parallel.shard_bounds(n_lights, world))
    subs, rows, pipes = [], [], []
    for r in range(world):
        sub, rws, _ = parallel.shard_scene(full, r, world, per_tree)
        subs.append(sub); rows.append(rws)
        pipes.append(bb.VisibilityPipeline(sub, device=r, world_size=world, rank=r, max_lights=max_lights))
    try:
        abi.p2p_link([p.ctx for p in pipes])
        world_o = OracleWorld(full)
        V = len(full.cameras)
        ranges = parallel.shard_bounds(n_lights, world)
        cap = ((max(1, max_lights) + 31) // 32) * 32
        for f in range(3):
            if f:
                scenes.advance_cameras(full)
                full_rows, _ = scenes.mutate_roots(full, f)
                world_o.tchanged[full_rows] = 1
                for r in range(world):
                    mine = np.isin(full_rows, rows[r])
                    local = np.searchsorted(rows[r], full_rows[mine]).astype(np.uint32)
                    subs[r].trs[local] = full.trs[full_rows[mine]]
                    pipes[r].ctx.upload_transforms_scattered(local, subs[r].trs[local])
            for p in pipes:
                p.update_views()
            for p in pipes:                                  # one host thread, asynchronous launches on both devices
                p.ctx.run(bb.STAGE_ALL)
            stats = [p.read_feedback() for p in pipes]
            planes = np.stack([np.ctypeslib.as_array(vw.half_spaces).reshape(6, 4).copy() for vw in pipes[0].views])
            _, _, lists, cl = world_o.frame(planes)
            for v in range(V):
                merged = parallel.merge_visible_lists([subs[r].entity_bits[pipes[r].ctx.download_visible(v)] for r in range(world)])
                assert merged.tolist() == full.entity_bits[lists[v]].tolist(), f"frame {f} view {v}: merged visible lists"
                out, off, idx = cl[v]
                nc = out.dims[0] * out.dims[1] * out.dims[2]
                for r in range(world):
                    goff, gidx = pipes[r].ctx.download_clusters(v)
                    assert (goff[:nc + 1] == off).all(), f"frame {f} view {v} device {r}: cluster offsets"
                    assert parallel.global_light_ordinal(gidx, cap, ranges).tolist() == idx.tolist()
                    assert stats[r].cluster_index_count[v] == out.total_index_count
                    assert np.float32(stats[r].cluster_farthest_z[v]).view(np.uint32) == np.float32(out.farthest_z).view(np.uint32)
    finally:
        for p in pipes:
            p.close()
