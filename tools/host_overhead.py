"""Measures host-side enqueue cost of the per-frame calls (GPU box).  Diagnostic only."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bevy_b200 as bb
from bevy_b200 import scenes

sc = scenes.forest()
pipe = bb.VisibilityPipeline(sc)
ctx = pipe.ctx
_s = torch.cuda.Stream(); torch.cuda.set_stream(_s); ctx.set_stream(_s.cuda_stream)
pipe.run_frame(); pipe.read_feedback()
rows, trs = scenes.mutate_roots(sc, 1)
rows_d = torch.from_numpy(rows.astype(np.int32)).cuda(); trs_d = torch.from_numpy(trs).cuda()
pipe.update_views()
slot = ctx.record_frame_constants()
ctx.use_recorded_frame_constants(slot)
n = len(rows)

def timeit(name, fn, iters=200):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:40s} host {1e6*(t1-t0)/iters:7.1f} us/iter   incl. drain {1e6*(t2-t0)/iters:7.1f} us/iter")

timeit("upload_transforms_scattered_raw", lambda: ctx.upload_transforms_scattered_raw(n, rows_d.data_ptr(), trs_d.data_ptr()))
timeit("use_recorded_frame_constants", lambda: ctx.use_recorded_frame_constants(slot))
timeit("run(ALL)", lambda: ctx.run(bb.STAGE_ALL))
timeit("run(PROPAGATE|CULL)", lambda: ctx.run(bb.STAGE_PROPAGATE | bb.STAGE_CULL))
timeit("run(PROPAGATE)", lambda: ctx.run(bb.STAGE_PROPAGATE))
timeit("run(CULL)", lambda: ctx.run(bb.STAGE_CULL))
timeit("run(CLUSTER)", lambda: ctx.run(bb.STAGE_CLUSTER))
def frame():
    ctx.upload_transforms_scattered_raw(n, rows_d.data_ptr(), trs_d.data_ptr()); ctx.use_recorded_frame_constants(slot); ctx.run(bb.STAGE_ALL)
timeit("full value_step", frame, 2000)
timeit("full value_step (100, below queue depth)", frame, 100)
