"""bench.py's CPU legs run without a GPU: the reference arm (the oracle timed on the host cores) must produce the contract's JSON line.
(Round 2 nearly shipped a reference arm that died on a missing binding -- nothing on the CPU side exercised it.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torchrun exports: the arm must override it
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "frames_per_sec" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    # the reference's own GLSL compiled for the CPU where that library exists (oracle/_ref), the restated oracle otherwise
    from oracle import glsl_ref
    assert line["cpu_baseline"]["kind"] == ("reference" if glsl_ref.available() else "port")
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(line["config"]) == {"workload", "sampler", "grid", "frame", "triangles", "parallelism", "frames_in_flight", "l2"}
    assert line["gpu_launches"] == 0


def test_reference_arm_non_root_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--config", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_frames_in_flight_rule_and_config_keys():
    """--frames-in-flight 0 = automatic: two pipelines for the reference's scenes, one for the large synthetic ones; NCCL exchange: always one"""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    a = argparse.Namespace(frames_in_flight=0)
    assert bench.frames_in_flight(a, 1112, 1, "p2p") == 2 and bench.frames_in_flight(a, 1112, 8, "p2p") == 2
    assert bench.frames_in_flight(a, 998412, 1, "p2p") == 1 and bench.frames_in_flight(a, 1112, 2, "nccl") == 1
    a.frames_in_flight = 3
    assert bench.frames_in_flight(a, 998412, 4, "p2p") == 3
    c1 = bench.config_dict(bench.CONFIGS[2], 1, 1, "p2p", 1112, 2)
    c8 = bench.config_dict(bench.CONFIGS[2], 8, 1, "p2p", 1112, 2)
    assert set(c1) == set(c8) and "every rank voxelizes the whole scene" in c8["parallelism"]
    assert "z-slab" in bench.config_dict(bench.CONFIGS[4], 8, 1, "p2p", 998412, 1)["parallelism"]


def test_screen_tile_lattice_is_a_balanced_partition():
    """capi.screen_tile_owner (= screen_tile_owner in csrc/vct_internal.cuh): every 32x32 tile has exactly one owner, the ranks' shares of a
    frame are balanced, and neighbours in a row AND in a column belong to different ranks"""
    import numpy as np
    sys.path.insert(0, ROOT)
    from voxel_cone_tracing_b200 import capi
    for W, H in ((1920, 1080), (7680, 4320), (320, 200)):
        ty, tx = np.meshgrid(np.arange((H + 31) // 32), np.arange((W + 31) // 32), indexing="ij")
        for n in (1, 2, 3, 4, 5, 6, 8, 15):
            own = capi.screen_tile_owner(tx, ty, n)
            assert own.min() >= 0 and own.max() < n
            counts = np.bincount(own.ravel(), minlength=n)
            assert counts.sum() == own.size
            if own.size >= 64 * n:
                assert counts.max() - counts.min() <= 0.05 * own.size / n + 2
            if n >= 2:
                assert (own[:, 1:] != own[:, :-1]).all() and (own[1:, :] != own[:-1, :]).all()
