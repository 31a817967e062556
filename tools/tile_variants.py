"""Device time of the tile pass alone (upload + run(PROPAGATE|CULL), bench workload) for the kernel variant selected by the
environment (B200VIS_TILE_KERNEL, B200VIS_WARP_VARIANT, ...).  GPU box, diagnostic: prints one line."""
import os, sys
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


def timeit(fn, iters=500):
    for i in range(30):
        fn(i)
    ctx.join(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(iters):
        fn(i)
    ctx.join(); e1.record(s); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


up = lambda i: ctx.upload_transforms_scattered_raw(n, rows_d.data_ptr(), frames[i % 8].data_ptr())   # noqa: E731
a = timeit(lambda i: (up(i), ctx.run(bb.STAGE_PROPAGATE | bb.STAGE_CULL)))
b = timeit(lambda i: (up(i), ctx.run(bb.STAGE_ALL)))
c = timeit(lambda i: ctx.run(bb.STAGE_CLUSTER))
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("B200VIS_"))
print(f"VARIANT [{tag}] prop+cull+expand {a:.1f} us  all(pipelined) {b:.1f} us  cluster-only {c:.1f} us", flush=True)
