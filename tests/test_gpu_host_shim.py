"""The C ABI without Python: tests/host_shim.c performs the Rust shim's start-up and per-frame sequence through
include/b200vis.h on Bevy-native column layouts (Transform 48 B, Affine3A 64 B, Aabb 32 B, Entity 8 B), with the GPU writing
its results straight into the "ECS" columns, and checks every frame against the CPU oracle."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_host_shim(out):
    sys.path.insert(0, ROOT)
    import oracle
    oracle.build()
    cmd = ["gcc", "-O2", "-std=gnu11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host_shim.c"), "-o", out,
           "-L" + os.path.join(ROOT, "bevy_b200"), "-lb200vis", "-L" + os.path.join(ROOT, "oracle"), "-lbevy_oracle", "-lm",
           "-Wl,-rpath," + os.path.join(ROOT, "bevy_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_header_is_plain_c_and_the_harness_links(tmp_path):
    """No GPU needed: include/b200vis.h compiles as C11 with -Wall -Wextra -Werror, every entry point the harness uses
    resolves against libb200vis.so, and without a CUDA device the library refuses to run (no CPU fallback)."""
    exe = str(tmp_path / "host_shim")
    build_host_shim(exe)
    import torch
    if not torch.cuda.is_available():
        res = subprocess.run([exe, "4", "3", "1"], capture_output=True, text=True, timeout=120)
        assert res.returncode == 3 and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_host_shim_sequence_matches_the_oracle(tmp_path):
    exe = str(tmp_path / "host_shim")
    build_host_shim(exe)
    res = subprocess.run([exe, "400", "6", "4"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "HOST_SHIM OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    stats = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert stats["entities"] == 400 * 63 + 48 and stats["step_ms_per_frame"] > 0


@pytest.mark.gpu
def test_host_shim_at_bench_scale_reports_host_costs(tmp_path):
    """1M entities (3922 trees of 255): the same sequence; prints what the host side costs per frame."""
    exe = str(tmp_path / "host_shim")
    build_host_shim(exe)
    res = subprocess.run([exe, "3922", "8", "3"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "HOST_SHIM OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    print(res.stdout[-600:])
