"""The benchmark inputs are pinned to what the REFERENCE's own loader (vendored tinyobjloader v1.1.0,
called as in src/renderer.cpp:417) reads from its assets: tests/golden/tinyobj_streams.json was produced
by oracle/_ref/tinyobj_dump (tools/make_golden.py).  Here the same digests are recomputed from the
committed fixtures and, when the reference tree is present, from our own OBJ/MTL reader."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import fixture_stream, oracle_digests, tinyobj_stream  # noqa: E402

from voxel_cone_tracing_b200 import scene as S  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "tinyobj_streams.json")))
PAIRS = [("CornellBox-Glossy.obj", "cornell_glossy.vctmesh"), ("suzanne.obj", "suzanne.vctmesh")]


@pytest.mark.parametrize("obj,fixture", PAIRS)
def test_fixture_matches_reference_loader_digest(obj, fixture):
    mesh = S.load_vctmesh(os.path.join(S.ASSET_DIR, fixture))
    hv, hm, n, k = fixture_stream(mesh)
    g = GOLD[obj]
    assert (n, k) == (g["n_indices"], g["n_materials"])
    assert hv == g["vertex_stream_sha256"]
    assert hm == g["materials_sha256"]


@pytest.mark.parametrize("obj,fixture", PAIRS)
def test_own_obj_reader_matches_tinyobj(obj, fixture):
    src = os.path.join("/root/reference/assets", obj)
    if not os.path.exists(src):
        pytest.skip("reference tree not present on this box")
    mesh = S.load_obj(src)
    hv, hm, n, k = fixture_stream(mesh)
    assert hv == GOLD[obj]["vertex_stream_sha256"] and hm == GOLD[obj]["materials_sha256"]
    dump = os.path.join(ROOT, "oracle", "_ref", "tinyobj_dump")
    if os.path.exists(dump):
        assert tinyobj_stream(src)[:2] == (hv, hm)      # live run of the reference's loader
    fx = S.load_vctmesh(os.path.join(S.ASSET_DIR, fixture))
    assert np.array_equal(fx.verts, mesh.verts) and np.array_equal(fx.indices, mesh.indices) and fx.ranges == mesh.ranges


def test_reference_scene_constants():
    sc = S.cornell_scene(with_suzanne=True)
    assert sc.n_triangles == 2080 and len(sc.draws) == 9 and len(sc.lights) == 1 and sc.cube_size == 3.0
    light = sc.materials[8]                       # 1 + mtl order: ... light is the 8th material
    assert np.allclose(light["emission"], 10.0) and light["illum"] == 2
    suz = sc.materials[sc.draws[-1]["material"]]  # src/main.cpp:88-97
    assert suz["illum"] == 4 and suz["shininess"] == 1000 and np.allclose(suz["emission"], (0, 0, 0.25))
    assert np.allclose(sc.draws[-1]["model"].reshape(4, 4)[3], (0, 1.1, -0.5, 1))


def test_oracle_regression_digests():
    """Guards the oracle against accidental edits (self-generated, not a reference pin)."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_regression.json")))
    now = oracle_digests()
    for k, v in gold.items():
        if k.startswith("frame_mean_rgba"):
            assert np.allclose(now[k], v, atol=0.05)
        elif k.startswith("samples"):
            assert abs(now[k] - v) <= 1e-3 * v
        else:
            assert now[k] == v, k


def test_synthetic_scene_is_seeded():
    a = S.synthetic_scene(30_000, 0x5EED0001)
    b = S.synthetic_scene(30_000, 0x5EED0001)
    c = S.synthetic_scene(30_000, 0x5EED0002)
    assert a.n_triangles == 12 + 5 * 5120
    assert np.array_equal(a.draws, b.draws) and not np.array_equal(a.draws, c.draws)
