"""torchrun entry: N ranks, one GPU each; every rank runs the CUDA path on its shard, the cluster slabs are
all-gathered over NCCL, and rank 0 checks the gathered cluster lists + merged visible lists against the
single-process CPU oracle on the whole scene (bit exact)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import bevy_b200 as bb                      # noqa: E402
from bevy_b200 import parallel, scenes      # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    per_tree, n_trees, n_lights = 63, 41, 37
    full = scenes.forest(n_trees=n_trees, levels=6, n_lights=n_lights)
    sub, rows, (l_lo, l_hi) = parallel.shard_scene(full, rank, world, per_tree)
    max_lights = max(hi - lo for lo, hi in parallel.shard_bounds(n_lights, world))
    # every rank is created with the same light capacity so the slabs have one size
    pipe = bb.VisibilityPipeline(sub, device=local, world_size=world, rank=rank, max_lights=max_lights)
    ctx = pipe.ctx
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    p2p = "--p2p" in sys.argv                # peer stores over NVLink (CUDA IPC), no collective on the data path
    builtin = "--builtin" in sys.argv or p2p   # the library's own exchange instead of the host's collective
    if p2p:
        mine = torch.from_numpy(ctx.p2p_export()).to(dev)
        handles = torch.zeros((world, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(handles.view(-1), mine)
        ctx.p2p_import(handles.cpu().numpy())
        dist.barrier()                       # nobody pushes before everybody has mapped everybody
    elif builtin:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.from_numpy(bb.Context.comm_unique_id()))
        dist.broadcast(uid, 0)
        ctx.comm_init(uid.cpu().numpy())
    else:
        slab = ctx.cluster_exchange_bytes()
        send = torch.zeros(slab // 4, dtype=torch.int32, device=dev)
        recv = torch.zeros(world * slab // 4, dtype=torch.int32, device=dev)
        ctx.set_cluster_exchange_buffers(send.data_ptr(), recv.data_ptr())
    V = len(full.cameras)
    ok = True
    if rank == 0:
        import oracle as orc
        from parity import OracleWorld
        world_o = OracleWorld(full)
    for f in range(3):
        if f:
            scenes.advance_cameras(full)            # cameras are shared objects between full and sub
            full_rows, _ = scenes.mutate_roots(full, f)     # ONE animation, defined on the full scene ...
            mine = np.isin(full_rows, rows)                 # ... of which this rank uploads its own roots
            local = np.searchsorted(rows, full_rows[mine]).astype(np.uint32)
            sub.trs[local] = full.trs[full_rows[mine]]
            ctx.upload_transforms_scattered(local, sub.trs[local])
        pipe.update_views()
        if builtin:
            ctx.run(bb.STAGE_ALL)
        else:
            ctx.run(bb.STAGE_PROPAGATE | bb.STAGE_CULL | bb.STAGE_CLUSTER_ASSIGN)
            with torch.cuda.stream(torch.cuda.ExternalStream(ctx.tail_stream(), device=dev)):
                parallel.all_gather_slabs(recv, send)          # on the stream the frame's tail runs on
            ctx.run(bb.STAGE_CLUSTER_LISTS)
        stats = ctx.download_frame_stats()
        # Clusters::last_frame_* come out of the list kernel identical on every rank: the index count is the total of the
        # gathered bit matrix, the farthest z is the max over the trailers that travelled with the slabs -- no reduction here
        far = np.array([stats.cluster_farthest_z[v] for v in range(V)], np.float32)
        cnt = np.array([stats.cluster_index_count[v] for v in range(V)], np.int64)
        chk = torch.from_numpy(np.concatenate([far.view(np.int32).astype(np.int64), cnt])).to(dev)
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        ok &= bool((lo_ == hi_).all().item())
        for v in range(V):
            fb = pipe.feedback[v]
            fb.has_farthest_z, fb.farthest_z, fb.has_index_count, fb.index_count = 1, float(far[v]), 1, int(cnt[v])
        vis_local = [sub.entity_bits[ctx.download_visible(v)] for v in range(V)]
        gathered = [None] * world
        dist.all_gather_object(gathered, [x.tolist() for x in vis_local])
        clusters = [ctx.download_clusters(v) for v in range(V)]
        if rank == 0:
            if f:
                world_o.tchanged[full_rows] = 1             # the oracle world holds the FULL scene
            planes = np.stack([np.ctypeslib.as_array(vw.half_spaces).reshape(6, 4).copy() for vw in pipe.views])
            _, _, lists, cl = world_o.frame(planes)
            ranges = parallel.shard_bounds(n_lights, world)
            cap = ((max(1, max_lights) + 31) // 32) * 32
            for v in range(V):
                merged = parallel.merge_visible_lists([g[v] for g in gathered])
                ok &= merged.tolist() == full.entity_bits[lists[v]].tolist()
                out, off, idx = cl[v]
                goff, gidx = clusters[v]
                nc = out.dims[0] * out.dims[1] * out.dims[2]
                ok &= bool((goff[:nc + 1] == off).all())
                glob = parallel.global_light_ordinal(gidx, cap, ranges)
                ok &= glob.tolist() == idx.tolist()
                ok &= int(cnt[v]) == out.total_index_count
                ok &= np.float32(far[v]).view(np.uint32) == np.float32(out.farthest_z).view(np.uint32)
            print(f"frame {f}: {'OK' if ok else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    pipe.close()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
