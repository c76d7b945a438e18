"""Multi-rank path (coordinate-range sharding, stitch, all-reduced reductions, z-slab halo
exchange): world_size-2 and -3 jobs over gloo on CPU with the kernel-logic simulator, and
(-m gpu, only when the box has >= 2 GPUs) over NCCL with the CUDA library."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(backend, world):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), backend],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=900)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out[-3000:]}"
    assert any("DIST_OK" in o for o in outs)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_path_gloo(world):
    import helpers
    helpers.build_sim()
    _launch("gloo-sim", world)


@pytest.mark.gpu
def test_sharded_path_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _launch("nccl-cuda", min(n, 4))
