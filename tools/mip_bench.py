"""Mip stage micro-benchmark (SURVEY 8(d) micro-inputs): dense builds -- every tile read and written -- of a uniform-random u32
grid, an opaque grid (a quarter of the channels are exact ties: worst case of the integer arithmetic) and the voxelized scene,
against the HBM roofline 7.4286 R^3 bytes / measured copy bandwidth.  CUDA events on the library's stream.

    python tools/mip_bench.py [R ...]        (default 256)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxel_cone_tracing_b200 import capi, scene as S  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_mip(dev, g, stream, reps=20):
    L = dev.L
    for _ in range(3):
        capi.check(L.vct_mipmap(dev.h, g.h))
    dev.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        capi.check(L.vct_mipmap(dev.h, g.h))
    e1.record(stream)
    dev.sync()
    return e0.elapsed_time(e1) * 1e3 / reps   # us


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256]
    dev = capi.Device(0)
    stream = torch.cuda.ExternalStream(int(dev.L.vct_device_stream(dev.h)))
    peak = peak_gbs()
    rng = np.random.default_rng(1)
    for R in sizes:
        levels = 7
        g = capi.Grid(dev, R, levels)
        alg = 7.4286 * R ** 3
        out = {"R": R, "levels": levels, "algorithmic_MB": alg / 1e6, "hbm_peak_gbs": peak, "ideal_us": alg / peak / 1e3}
        for kind in os.environ.get("KINDS", "random,opaque,scene").split(","):
            if kind == "scene":
                sc = S.cornell_scene()
                p = capi.DeviceScene(dev, sc)
                capi.check(dev.L.vct_voxelize_reserve(dev.h, 1 << 23))
                g.clear()
                capi.check(dev.L.vct_voxelize(dev.h, p.h, g.h, 0, R))
                dev.sync()
            else:
                # built slab by slab: a 1024^3 grid is 4 GiB
                base = np.empty((R, R, R), np.uint32)
                for z in range(0, R, 64):
                    blk = rng.integers(0, 2 ** 32, (min(64, R), R, R), dtype=np.uint64).astype(np.uint32)
                    if kind == "opaque":
                        blk |= np.uint32(0xFF000000)
                    base[z:z + 64] = blk
                g.upload_base(base)
                del base
            dev.debug_set(capi.DEBUG_MIP_DENSE, 1)
            us = time_mip(dev, g, stream)
            dev.debug_set(capi.DEBUG_MIP_DENSE, 0)
            out[f"dense_{kind}_us"] = round(us, 2)
            out[f"dense_{kind}_frac"] = round(alg / (us * 1e-6) / 1e9 / peak, 3)
            if kind == "scene":
                out["sparse_scene_us"] = round(time_mip(dev, g, stream), 2)
                p.close()
        print(json.dumps(out), flush=True)
        g.close()
    dev.close()


if __name__ == "__main__":
    main()
