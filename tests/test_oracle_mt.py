"""The multithreaded CPU baseline (oracle/bevy_oracle_mt.c, used by bench.py) is bit-identical to the serial oracle."""
import numpy as np

import oracle as orc
from bevy_b200 import scenes
from parity import OracleWorld


def test_mt_baseline_matches_serial_oracle():
    a, b = scenes.forest(n_trees=60, levels=7, n_lights=12), scenes.forest(n_trees=60, levels=7, n_lights=12)
    wa, wb = OracleWorld(a), OracleWorld(b)
    for f in range(3):
        if f:
            for sc, w in ((a, wa), (b, wb)):
                scenes.advance_cameras(sc)
                rows, _ = scenes.mutate_roots(sc, f)
                w.tchanged[rows[::3]] = 1          # only a third of the trees are dirty: exercises the static skip
        planes = np.stack([orc.compute_frustum(orc.perspective(c.fov, c.aspect, c.near), c.gt, c.far) for c in a.cameras])
        ra = wa.frame(planes, mt=False)
        rb = wb.frame(planes, mt=True)
        assert (wa.gt.view(np.uint32) == wb.gt.view(np.uint32)).all()
        assert (wa.vv == wb.vv).all()
        assert (ra[0] == rb[0]).all() and (ra[1] == rb[1]).all()
        for la, lb in zip(ra[2], rb[2]):
            assert (la == lb).all()
