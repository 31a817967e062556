"""GPU parity tests proper: CUDA path through the C ABI vs the CPU oracle (bit exact)."""
import numpy as np
import pytest

import bevy_b200 as bb
from bevy_b200 import scenes

from parity import run_parity

pytestmark = pytest.mark.gpu


def test_flat_many_cubes_small():
    run_parity(scenes.many_cubes(20_000, n_lights=64, light_range=(0.3, 8.0)), frames=3)


def test_forest_small_static_opt_enabled():
    run_parity(scenes.forest(n_trees=200, levels=8, n_lights=64), frames=4, static_opt=True)


def test_forest_small_static_opt_disabled():
    run_parity(scenes.forest(n_trees=120, levels=6, n_lights=32, seed=7), frames=3, static_opt=False)


def test_propagate_bench_scene_multi_pass_plan():
    # config #1: 1077-node trees do not fit one 256-row tile -> parents in other tiles, several passes
    run_parity(scenes.propagate_bench_scene(), frames=3, cluster=False)


def test_config2_many_cubes_160k():
    run_parity(scenes.many_cubes(160_000), frames=2, cluster=False)
