"""How much does the GPU gain from running two independent frame streams concurrently? (diagnostic)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bevy_b200 as bb
from bevy_b200 import scenes

def make(seed):
    sc = scenes.forest(seed=seed)
    pipe = bb.VisibilityPipeline(sc)
    s = torch.cuda.Stream()
    pipe.ctx.set_stream(s.cuda_stream)
    pipe.run_frame(); pipe.read_feedback()
    frames = [torch.from_numpy(scenes.mutate_roots(sc, f + 1)[1]).cuda() for f in range(8)]
    rows = torch.from_numpy(sc.roots.astype(np.int32)).cuda()
    pipe.update_views_fast()
    slot = pipe.ctx.record_frame_constants()
    pipe.ctx.use_recorded_frame_constants(slot)
    return pipe, s, frames, rows

def run(pipes, iters=1500):
    for i in range(50):
        for (p, s, fr, rows) in pipes:
            p.ctx.upload_transforms_scattered_raw(len(rows), rows.data_ptr(), fr[i % 8].data_ptr()); p.ctx.run(bb.STAGE_ALL)
    for (p, s, fr, rows) in pipes: p.ctx.join()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(iters):
        for (p, s, fr, rows) in pipes:
            p.ctx.upload_transforms_scattered_raw(len(rows), rows.data_ptr(), fr[i % 8].data_ptr()); p.ctx.run(bb.STAGE_ALL)
    for (p, s, fr, rows) in pipes: p.ctx.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = sum(p.scene.n for (p, _, _, _) in pipes)
    print(f"{len(pipes)} context(s): {1e6*dt/iters:7.1f} us per round, {n*iters/dt/1e9:6.2f} G entities/s aggregate")

a = make(1)
run([a])
b = make(2)
run([a, b])
run([a])
