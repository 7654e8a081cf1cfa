#!/usr/bin/env python
"""The seeded random scenes of tests/edge_scenes.py (general model matrices, triangle soups reaching outside the cube, cameras inside the
geometry) through the reference's passes on Mesa llvmpipe (oracle/gl_ref) against the oracle: fragment counts, folded voxel grid, frames.

    python tools/gl_llvmpipe_random_scenes.py 0 24 > profiles/r02_gl_llvmpipe_random_scenes.txt
"""
import sys, time
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import orc, gl_ref
from voxel_cone_tracing_b200 import scene as S
from edge_scenes import fuzz_case, edge_scene, EDGE_KINDS
def one(name, sc, R, levels, W, H, cam, kw):
    view, proj = S.reference_camera(W / H, **cam)
    # voxelization fragments
    tri, xy, vox, col = gl_ref.voxelize_fragments(sc, R)
    base, st = orc.voxelize(sc, R)
    v = vox.astype(np.int64); inb = ((v>=0)&(v<R)).all(axis=1)
    grid = np.zeros((R,R,R), np.uint32)
    for i in np.where(inb)[0]:
        x,y,z = v[i]; grid[z,y,x] = orc.fold(int(grid[z,y,x]), col[i])
    dv = np.abs(grid.view(np.uint8).astype(int)-base.view(np.uint8).astype(int)).reshape(-1,4).max(axis=1)
    # direct-only frame (no fetch) and full frame with brilinear
    pyr = orc.mipmap(base, levels)
    g = orc.gbuffer(sc, view, proj, W, H)
    res=[]
    for label, k2, mode in (('direct', dict(enable_diffuse=0, enable_specular=0, enable_shadow=0, view_voxel_dir=7), 0), ('as drawn', kw, 1)):
        prm = orc.default_params(**k2)
        if prm.view_voxel_dir < 7 and label=='as drawn': res.append('debug view skipped'); continue
        orc.debug_set_lod_filter(mode)
        fr,_ = orc.trace(sc, view, g, pyr, prm)
        orc.debug_set_lod_filter(0)
        u8,_ = gl_ref.visualize(sc, view, proj, pyr, W, H, prm)
        d = np.abs(fr.view(np.uint8).reshape(H,W,4).astype(int)-u8.view(np.uint8).reshape(H,W,4).astype(int)).max(axis=2)
        res.append(f'{label}: bg {"=" if np.array_equal(g.tri_id==0xFFFFFFFF, u8==0xFF404026) else "DIFF %d" % ((g.tri_id==0xFFFFFFFF)!=(u8==0xFF404026)).sum()} max {d.max()} >1 {(d>1).sum()} >2 {(d>2).sum()}/{W*H}')
    print(f'{name}: tris {sc.n_triangles} R {R} frags gl {len(tri)} (in {inb.sum()}) oracle {st.fragments} (oob {st.fragments_oob}); occ {"=" if np.array_equal(grid!=0,base!=0) else "DIFF"} cnt {"=" if not ((grid^base)&0x01010101).any() else "DIFF %d" % (((grid^base)&0x01010101)!=0).sum()} vox-diff {(dv>0).sum()}/{st.occupied} max {dv.max()} | ' + ' | '.join(res), flush=True)
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    sc, R, levels, W, H, cam, kw = fuzz_case(seed, False)
    try: one(f'fuzz {seed}', sc, R, levels, W, H, cam, kw)
    except Exception as ex: print('fuzz', seed, 'EXC', repr(ex)[:300])
