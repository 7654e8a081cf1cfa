"""Voxelizer: ordered (reference, bit-exact) against fixed-point (order-independent integer mean) accumulation -- stage time and the
per-channel difference between the two grids, configs 2, 4, 5.   python tools/accum_modes.py [config ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from voxel_cone_tracing_b200 import capi  # noqa: E402


def main():
    for cid in [int(a) for a in sys.argv[1:]] or [2, 4, 5]:
        cfg = bench.CONFIGS[cid]
        sc = bench.build_scene(cfg)
        R = cfg["R"]
        p = capi.Pipeline(sc, R, 64, 64, 7, reserve=max(1 << 20, 24 * sc.n_triangles))
        stream = torch.cuda.ExternalStream(int(p.dev.L.vct_device_stream(p.dev.h)))
        out = {"config": cid, "grid": R, "triangles": sc.n_triangles}
        grids = {}
        for name, mode in (("ordered", capi.ACCUM_ORDERED), ("fixed_point", capi.ACCUM_FIXED_POINT)):
            p.dev.set_accum_mode(mode)
            for _ in range(3):
                p.clear(); p.voxelize()
            p.sync()
            st = p.voxel_stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record(stream)
            for _ in range(n):
                p.clear(); p.voxelize()
            e1.record(stream)
            p.sync()
            us = e0.elapsed_time(e1) * 1e3 / n
            out[name + "_clear_plus_voxelize_us"] = round(us, 1)
            out[name + "_gfrag_per_s"] = round(st.fragments / us / 1e3, 2)
            out["fragments"], out["occupied"], out["max_per_voxel"] = int(st.fragments), int(st.occupied), int(st.max_per_voxel)
            grids[name] = p.grid.download(0)
        occ = grids["ordered"] != 0
        a = grids["ordered"][occ].view(np.uint8).astype(np.int16)
        b = grids["fixed_point"][occ].view(np.uint8).astype(np.int16)
        d = np.abs(a - b)
        out["max_abs_diff_of_255"], out["mean_abs_diff_of_255"] = int(d.max()), round(float(d.mean()), 3)
        out["same_occupancy"] = bool(np.array_equal(occ, grids["fixed_point"] != 0))
        print(json.dumps(out), flush=True)
        p.close()


if __name__ == "__main__":
    main()
