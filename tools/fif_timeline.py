"""Timeline of two frames in flight on N GPUs (run under torch.distributed.run): event times of the last frame of pipeline A and of
pipeline B, relative to the start of A's, on every rank.  Shows where the period of the alternating loop goes (front half stretched by
the other frame's cone kernel, flag waits, the shared trace stream).   python -m torch.distributed.run --nproc-per-node N tools/fif_timeline.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxel_cone_tracing_b200 import capi, scene as S  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R, W, H = 256, 1920, 1080
sc = S.cornell_scene()
view, proj = S.reference_camera(W / H)
prm = capi.default_params(sampler=1, tile_rank=rank, tile_nranks=world)
pipes = []
for k in range(2):
    p = capi.Pipeline(sc, R, W, H, ordinal=local)
    p.dev.debug_set(capi.DEBUG_TRACE_LOW_PRIORITY, int(os.environ.get("LOW", "1")))
    if world > 1:
        hs = [None] * world
        dist.all_gather_object(hs, p.peer_export())
        p.peer_connect(rank, world, hs, frame_root=0)
    pipes.append(p)
if world > 1:
    dist.barrier()
n = 41   # odd: the last frame is A's, the one before B's
for i in range(n):
    pipes[i & 1].render_frame(view, proj, prm)
for p in pipes:
    p.sync()
names = ("start", "clear", "vox", "mip", "front", "end", "cone0", "cone1")
rows = {}
for tag, p in (("B (frame n-2)", pipes[1]), ("A (frame n-1)", pipes[0])):
    t = np.zeros(8, np.float32)
    capi.check(p.dev.L.vct_debug_frame_events(p.dev.h, pipes[1].dev.h, t.ctypes.data_as(C.POINTER(C.c_float))))
    rows[tag] = {k: round(float(v) * 1e3, 1) for k, v in zip(names, t)}
out = [None] * world
if world > 1:
    dist.all_gather_object(out, rows)
else:
    out = [rows]
if rank == 0:
    print("us relative to the start of B's last frame; order of a frame: start < clear < vox < mip < front < cone0 < cone1 < end")
    for r, rw in enumerate(out):
        for tag, d in rw.items():
            print(f"rank {r} {tag}: " + "  ".join(f"{k} {d[k]:8.1f}" for k in ("start", "clear", "vox", "mip", "front", "cone0", "cone1", "end")))
if world > 1:
    for p in pipes:
        p.peer_check()
    dist.barrier()
    for p in pipes:
        p.peer_disconnect()
    dist.barrier()
for p in pipes:
    p.close()
if world > 1:
    dist.destroy_process_group()
