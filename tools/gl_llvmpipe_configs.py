#!/usr/bin/env python
"""BASELINE configs 1 and 2 at full size through the reference's three passes on Mesa llvmpipe (oracle/gl_ref) against the oracle:
profiles/r02_gl_llvmpipe_configs.jsonl, summarised in profiles/r02_gl_llvmpipe_parity.md.  CPU only, about 4 minutes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc, gl_ref
from voxel_cone_tracing_b200 import scene as S
def run(name, suz, R, W, H, out):
    sc = S.cornell_scene(with_suzanne=suz)
    view, proj = S.reference_camera(W / H)
    t=time.time(); tri, xy, vox, col = gl_ref.voxelize_fragments(sc, R); tv=time.time()-t
    base, st = orc.voxelize(sc, R)
    grid = np.zeros((R,R,R), np.uint32); v = vox.astype(np.int64)
    for i in range(len(tri)):
        x,y,z = v[i]; grid[z,y,x] = orc.fold(int(grid[z,y,x]), col[i])
    dv = np.abs(grid.view(np.uint8).astype(int)-base.view(np.uint8).astype(int)).reshape(-1,4).max(axis=1)
    pyr = orc.mipmap(base, 7)
    t=time.time(); glm = gl_ref.mip_chain(base, 7); tm=time.time()-t
    n_m = sum(pyr.levels[d][l].size for d in range(6) for l in range(1,7))
    off_plain = sum(int((pyr.levels[d][l]!=glm[d][l]).sum()) for d in range(6) for l in range(1,7))
    orc.debug_set_unorm_unpack(1); orc.debug_set_mip_balanced_sum(1)
    pm = orc.mipmap(base, 7)
    orc.debug_set_unorm_unpack(0); orc.debug_set_mip_balanced_sum(0)
    off_model = sum(int((pm.levels[d][l]!=glm[d][l]).sum()) for d in range(6) for l in range(1,7))
    g = orc.gbuffer(sc, view, proj, W, H)
    t=time.time(); u8, f32 = gl_ref.visualize(sc, view, proj, pyr, W, H); tf=time.time()-t
    res = {}
    for mode in (0,1):
        orc.debug_set_lod_filter(mode)
        fr,_ = orc.trace(sc, view, g, pyr)
        d = np.abs(fr.view(np.uint8).reshape(H,W,4).astype(int)-u8.view(np.uint8).reshape(H,W,4).astype(int)).max(axis=2)
        mse = ((fr.view(np.uint8).astype(float)-u8.view(np.uint8).astype(float))**2).mean()
        res[mode] = dict(max=int(d.max()), differing=int((d>0).sum()), over1=int((d>1).sum()), over2=int((d>2).sum()), psnr=float(10*np.log10(255**2/max(mse,1e-12))))
    orc.debug_set_lod_filter(0)
    r = dict(case=name, R=R, frame=[W,H], triangles=sc.n_triangles,
             voxelize=dict(fragments_gl=len(tri), fragments_oracle=int(st.fragments), occupied=int(st.occupied), occupancy_equal=bool(np.array_equal(grid!=0,base!=0)),
                           counts_equal=bool(not ((grid^base)&0x01010101).any()), voxels_differing=int((dv>0).sum()), max_channel_diff=int(dv.max()), gl_seconds=round(tv,1)),
             mip=dict(texels=n_m, differing_plain_rules=off_plain, differing_driver_model=off_model, gl_seconds=round(tm,1)),
             frame_vs_gl=dict(background_mask_equal=bool(np.array_equal(g.tri_id==0xFFFFFFFF, u8==0xFF404026)), pixels=W*H, rule_R7_linear=res[0], llvmpipe_brilinear_modelled=res[1], gl_seconds=round(tf,1)))
    print(json.dumps(r)); out.write(json.dumps(r)+'\n'); out.flush()
with open(os.path.join(ROOT, 'profiles', 'r02_gl_llvmpipe_configs.jsonl'), 'w') as out:
    run('config 1 (CornellBox-Glossy 128^3, 512x512)', False, 128, 512, 512, out)
    run('config 1 + Suzanne', True, 128, 512, 512, out)
    run('config 2 (CornellBox-Glossy 256^3, 1920x1080)', False, 256, 1920, 1080, out)
