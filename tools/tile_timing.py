"""Where a tile-kernel CTA spends its time (debug build, not the product library).

    cd bevy_b200/csrc && nvcc <the flags of bevy_b200/build.py> -DB200VIS_TILE_TIMING \
        -o ../../build/libb200vis_timing.so kernels.cu api.cu host_view.cpp
    B200VIS_LIB=build/libb200vis_timing.so python tools/tile_timing.py

Prints, averaged over the two-tile CTAs of one launch on the bench workload, the clock64 deltas between the phase
marks thread 0 leaves in k_propagate_cull_tma while it processes its SECOND tile (steady state): load wait, dirty
phase, every hierarchy level, cull, end barrier.
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevy_b200 as bb  # noqa: E402
from bevy_b200 import abi, scenes  # noqa: E402


def main():
    sc = scenes.forest(3922, 8, 256)
    pipe = bb.VisibilityPipeline(sc)
    ctx = pipe.ctx
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    for f in range(30):
        scenes.advance_cameras(sc)
        rows, trs = scenes.mutate_roots(sc, f + 1)
        ctx.upload_transforms_scattered(rows, trs)
        pipe.update_views_fast()
        ctx.run(bb.STAGE_ALL)
    ctx.join(); torch.cuda.synchronize()
    lib = abi.load_library()
    n = 2368
    buf = np.zeros((n, 16), np.uint64)
    rc = lib.b200vis_debug_tile_timing(ctypes.c_void_p(buf.ctypes.data), n)
    assert rc == 0, rc
    t = buf.astype(np.int64)
    ok = (t[:, 0] > 0) & (t[:, 14] > 0)
    two = ok & (t[:, 8] > 0)
    two = ok & (t[:, 2] > 0) & (t[:, 15] > 0)
    print("CTAs with marks:", int(ok.sum()), "two-tile CTAs:", int(two.sum()))

    def d(a, b, m):
        x = (t[m, b] - t[m, a])
        return f"{x.mean():9.0f} (p10 {np.percentile(x, 10):7.0f}  p90 {np.percentile(x, 90):7.0f})"
    print("second tile of a CTA (steady state), thread 0, cycles (clock64), mean (p10, p90):")
    print(" CTA start -> tile1 loaded        ", d(0, 2, two))
    print(" dirty phase (one CTA barrier)    ", d(2, 3, two))
    print(" local affine + level 0           ", d(3, 4, two))
    for lvl in range(1, 8):
        print(f" level {lvl} (barrier + work)          ", d(3 + lvl, 4 + lvl, two))
    print(" flags write-back -> walk done    ", d(11, 12, two))
    print(" cull                             ", d(12, 13, two))
    print(" end barrier                      ", d(13, 15, two))
    print(" tile1 total                      ", d(2, 15, two))
    print(" CTA lifetime                     ", d(0, 14, two))
    pipe.close()


if __name__ == "__main__":
    main()
