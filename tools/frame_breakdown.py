"""Device time per frame for sub-sets of the frame's work (GPU box, diagnostic)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bevy_b200 as bb
from bevy_b200 import scenes

sc = scenes.forest()
pipe = bb.VisibilityPipeline(sc)
ctx = pipe.ctx
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
pipe.run_frame(); pipe.read_feedback()
frames = []
for f in range(8):
    rows, trs = scenes.mutate_roots(sc, f + 1)
    frames.append(torch.from_numpy(trs).cuda())
rows_d = torch.from_numpy(sc.roots.astype(np.int32)).cuda()
pipe.update_views_fast()
slot = ctx.record_frame_constants()
ctx.use_recorded_frame_constants(slot)
n = len(sc.roots)

def timeit(name, fn, iters=1000):
    for i in range(50): fn(i)
    ctx.join(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(iters): fn(i)
    ctx.join(); e1.record(s); torch.cuda.synchronize()
    print(f"{name:46s} {1e3*e0.elapsed_time(e1)/iters:7.2f} us/frame")

up = lambda i: ctx.upload_transforms_scattered_raw(n, rows_d.data_ptr(), frames[i % 8].data_ptr())
timeit("upload + run(ALL)", lambda i: (up(i), ctx.run(bb.STAGE_ALL)))
timeit("upload + run(PROP|CULL)", lambda i: (up(i), ctx.run(bb.STAGE_PROPAGATE | bb.STAGE_CULL)))
timeit("upload + run(PROP)", lambda i: (up(i), ctx.run(bb.STAGE_PROPAGATE)))
timeit("upload only", lambda i: up(i))
timeit("run(ALL) static (nothing dirty)", lambda i: ctx.run(bb.STAGE_ALL))
timeit("run(CULL) only", lambda i: ctx.run(bb.STAGE_CULL))
timeit("run(CLUSTER) only", lambda i: ctx.run(bb.STAGE_CLUSTER))
ctx.set_static_transform_optimizations(False)
timeit("run(PROP) static_opt disabled (full recompute)", lambda i: ctx.run(bb.STAGE_PROPAGATE))
timeit("run(ALL) static_opt disabled", lambda i: ctx.run(bb.STAGE_ALL))
