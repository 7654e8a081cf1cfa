#!/usr/bin/env python
"""One-off: BASELINE config 5's scene (4 M synthetic triangles, 1024^3) voxelized by THE REFERENCE'S OWN GLSL on the CPU
(oracle/_ref/libvct_glsl_ref.so, 72.7 M fragments through voxelize.frag's compare-and-swap loop, six 4 GiB textures) and by the
oracle; levels 1..3 of the mip chain on the central 256^3 block.  Needs ~45 GB of host memory: too heavy for the CPU test suite,
so the result is kept under profiles/.

    python tools/glsl_ref_config5.py > profiles/r02_glsl_ref_config5.txt
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from oracle import glsl_ref as G  # noqa: E402
from oracle import orc  # noqa: E402

cfg = bench.CONFIGS[5]
sc = bench.build_scene(cfg)
R = cfg["R"]
print(f"scene: {sc.n_triangles} triangles, grid {R}^3")
t = time.time(); base, st = orc.voxelize(sc, R)
print(f"oracle        : {time.time() - t:6.1f} s  fragments {st.fragments}  occupied {st.occupied}  max per voxel {st.max_per_voxel}  wrapped voxels {st.wrapped_voxels}")
t = time.time(); tex, n = G.voxelize(sc, R, "rules")
print(f"reference GLSL: {time.time() - t:6.1f} s  fragments executed {n}")
for i in range(6):
    print(f"  texture {i}: {int((tex[i] != base).sum())} voxels differ from the oracle")
sub = np.ascontiguousarray(base[384:640, 384:640, 384:640])
del tex
po, pg = orc.mipmap(sub, 4), G.mipmap(sub, 4, "rules")
print("mip levels 1..3 of the central 256^3 block: " + ("identical" if all(np.array_equal(po.levels[d][l], pg.levels[d][l]) for d in range(6) for l in range(1, 4)) else "DIFFER"))
