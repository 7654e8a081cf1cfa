"""Worker of tests/test_multi_gpu.py (one process per rank, launched with torch.distributed.run, gloo for the plumbing).

mode "gpu":  every rank renders its share of K frames through the NVLink peer-memory exchange (vct_peer_*); rank 0 checks that the
             exchanged voxel grid and the merged frame are BIT-IDENTICAL to a single-GPU render of the same scene in its own process.
             Uses GPU `local_rank` when the box has that many GPUs, else all ranks share GPU 0 (CUDA IPC works within one device too).
mode "cpu":  the same partitioning (z-slabs, 32x32 tile ownership) exercised on the CPU with the oracle as the compute and gloo
             collectives as the exchange: merged slabs == full voxelization, merged tiles == full frame."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxel_cone_tracing_b200 import scene as S  # noqa: E402


def main():
    mode, out_path = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    sc = S.cornell_scene(with_suzanne=True, theta=0.3)
    R, W, H = 64, 320, 200
    view, proj = S.reference_camera(W / H)
    result = {"rank": rank, "ok": False}
    if mode == "cpu":
        from oracle import orc
        full_base, _ = orc.voxelize(sc, R)
        z0, z1 = rank * R // world, (rank + 1) * R // world
        part, _ = orc.voxelize(sc, R, z0, z1)
        assert not part[:z0].any() and not part[z1:].any()
        t = torch.from_numpy(part.astype(np.int64))
        dist.all_reduce(t)                      # slabs are disjoint: sum == merge
        merged = t.numpy().astype(np.uint32)
        pyr = orc.mipmap(merged, 7)
        g = orc.gbuffer(sc, view, proj, W, H)
        frame = np.zeros((H, W), np.uint32)
        orc.trace(sc, view, g, pyr, None, world, rank, frame)      # tiles t with t % world == rank
        ft = torch.from_numpy(frame.astype(np.int64))
        dist.all_reduce(ft)
        ref = orc.render_frame(sc, view, proj, R, W, H)
        result["ok"] = bool(np.array_equal(merged, full_base) and np.array_equal(ft.numpy().astype(np.uint32), ref["frame"]))
    else:
        from voxel_cone_tracing_b200 import capi
        ngpu = torch.cuda.device_count()
        ordinal = rank if ngpu >= world else 0
        prm = capi.default_params(sampler=int(os.environ.get("VCT_TEST_SAMPLER", "1")))
        n_frames = 5                            # a moving object: exercises the double buffering, the epoch flags and the sparse un-push of old voxels
        scenes = [S.cornell_scene(with_suzanne=True, theta=0.3 + 0.4 * k) for k in range(n_frames)]
        ref_frames, ref_bases, ref_l2 = [], [], None
        if rank == 0:   # single-GPU reference in this process
            p1 = capi.Pipeline(sc, R, W, H, ordinal=ordinal)
            for k in range(n_frames):
                p1.scene.upload(scenes[k])
                p1.render_frame(view, proj, prm)
                ref_frames.append(p1.target.frame().copy()); ref_bases.append(p1.grid.download(0)); ref_l2 = p1.grid.download(2, 3)
            p1.close()
        rb = torch.from_numpy(np.stack(ref_bases).astype(np.int64)) if rank == 0 else torch.zeros((n_frames, R, R, R), dtype=torch.int64)
        dist.broadcast(rb, 0)                   # every rank checks its gathered grid against the single-GPU grid, every frame
        ref_bases = rb.numpy().astype(np.uint32)
        # VCT_TEST_REPLICATE: 0 = z-slab voxelization + sparse voxel push, 1 = every rank voxelizes the whole scene (small-scene mode);
        # VCT_TEST_FIF: pipelines rendering alternate frames (frames in flight; their traces share one low-priority stream per GPU)
        replicate, fif = int(os.environ.get("VCT_TEST_REPLICATE", "0")), int(os.environ.get("VCT_TEST_FIF", "1"))
        pipes = []
        for _ in range(fif):
            pipe = capi.Pipeline(sc, R, W, H, ordinal=ordinal)
            pipe.dev.debug_set(capi.DEBUG_PEER_REPLICATE, replicate)
            if fif > 1:
                pipe.dev.debug_set(capi.DEBUG_TRACE_LOW_PRIORITY, 1)
            handles = [None] * world
            dist.all_gather_object(handles, pipe.peer_export())
            pipe.peer_connect(rank, world, handles, frame_root=0)
            pipes.append(pipe)
        dist.barrier()
        ok = True
        for k in range(n_frames):
            pipe = pipes[k % fif]
            pipe.scene.upload(scenes[k])
            pipe.render_frame(view, proj, prm)
            pipe.sync()
            dist.barrier()                      # the download below reads this rank's copy: every peer's pushes of frame k are complete (flag wait), nobody has started k+1
            if rank == 0:
                ok &= bool(np.array_equal(pipe.target.frame(), ref_frames[k]))
            ok &= bool(np.array_equal(pipe.grid.download(0), ref_bases[k]))
            dist.barrier()
        # the same frames again WITHOUT host synchronisation in between: ranks run ahead of each other as far as the flags allow
        for k in range(2 * n_frames + 1):
            pipe = pipes[k % fif]
            pipe.scene.upload(scenes[k % n_frames])
            pipe.render_frame(view, proj, prm)
        for q in pipes:
            q.sync()
        dist.barrier()
        ref_base = ref_bases[0]                 # the last frame rendered was scene 0, on `pipe`
        if rank == 0:
            ok &= bool(np.array_equal(pipe.target.frame(), ref_frames[0]))
            ref_l2 = None
        base = pipe.grid.download(0)
        bt = torch.from_numpy(base.astype(np.int64)); b0 = bt.clone()
        dist.broadcast(b0, 0)
        ok &= bool(torch.equal(bt, b0))         # every rank ends up with the same full grid
        if rank == 0:
            ok &= bool(np.array_equal(base, ref_base))
        for q in pipes:
            q.peer_check()
        dist.barrier()
        for q in pipes:
            q.peer_disconnect()
        dist.barrier()
        for q in pipes:
            q.close()
        result["ok"] = ok
        result["gpus"] = ngpu
    flags = [None] * world
    dist.all_gather_object(flags, result["ok"])
    if rank == 0:
        json.dump({"ok": all(flags), "per_rank": flags, **{k: v for k, v in result.items() if k not in ("ok",)}}, open(out_path, "w"))
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
