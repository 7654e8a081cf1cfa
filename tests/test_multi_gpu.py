"""Multi-rank tests (one process per rank under torch.distributed.run, rendezvous on 127.0.0.1).

CPU: world_size 2 and 3 with gloo -- the sharding logic (z-slabs, screen-tile ownership, merges) against the oracle.
GPU: world_size 2 -- the NVLink peer-memory exchange (vct_peer_export / connect, sparse voxel push fused into the
     resolve kernel, tile push fused into the shade kernel, epoch flags) must reproduce the single-GPU grid and frame
     bit for bit."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mp_worker.py")


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(mode: str, world: int, tmp_path, timeout: int, env_extra=None):
    out = tmp_path / f"{mode}{world}.json"
    env = dict(os.environ, OMP_NUM_THREADS="2", **(env_extra or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), WORKER, mode, str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.load(open(out))


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_logic_on_cpu_gloo(tmp_path, world):
    res = _launch("cpu", world, tmp_path, 600)
    assert res["ok"], res


@pytest.mark.gpu
@pytest.mark.parametrize("sampler,replicate,fif", [(0, 0, 1), (1, 0, 1), (1, 1, 1), (1, 0, 2), (1, 1, 2)])
def test_peer_exchange_two_ranks_bit_identical(tmp_path, sampler, replicate, fif):
    """sampler: software / texture-unit cone sampler; replicate: z-slab voxelization + voxel push (0) or small-scene mode, every rank voxelizes
    everything (1); fif: pipelines rendering alternate frames (frames in flight)"""
    res = _launch("gpu", 2, tmp_path, 600, {"VCT_TEST_SAMPLER": str(sampler), "VCT_TEST_REPLICATE": str(replicate), "VCT_TEST_FIF": str(fif)})
    assert res["ok"], res
