"""Multi-GPU bit-exactness under pytest (-m gpu; skipped below two GPUs): the slab (NCCL) and peer (NVLink stores) clipmap
builds equal the single-GPU build bit for bit over several frames with a moving camera and cadence, and row-sharded /
tile-interleaved cone tracing equals the full-image trace. The ranks are launched exactly as the driver launches
bench.py (torch.distributed.run, one process per GPU, rendezvous on 127.0.0.1); the checks themselves live in
tools/multigpu_check.py so that the same code also produces the timings under profiles/."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, *extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "multigpu_check.py"), *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines, f"multigpu_check failed (rc {p.returncode}):\n{p.stdout[-2000:]}\n{p.stderr[-4000:]}"
    return json.loads(lines[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_builds_and_traces_bit_exact_cornell(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _run(world, "--res", "64", "--frames", "4")
    assert r["world"] == world
    assert r["slab_build_bit_exact"] and r["peer_build_bit_exact"] and r["row_sharded_trace_bit_exact"], r


@pytest.mark.gpu
def test_sharded_builds_bit_exact_atrium_six_levels():
    world = 2
    if _gpu_count() < world:
        pytest.skip("needs 2 GPUs")
    r = _run(world, "--scene", "atrium", "--res", "128", "--frames", "3", "--width", "320", "--height", "184")
    assert r["slab_build_bit_exact"] and r["peer_build_bit_exact"] and r["row_sharded_trace_bit_exact"], r
