"""Differential test of the two OBJ / MTL readers (voxel_cone_tracing_b200/scene.py load_obj, host/obj_loader.cpp behind
Renderer::load_model) against the loader the reference uses (vendored tinyobjloader v1.1.0, src/renderer.cpp:417) on seeded
random files (tests/obj_fuzz.py): every syntax an exporter writes, and in every third case also what none would (numbers the
loader's own float parser half-accepts, `usemtl` in front of `o`, names with blanks, several / repeated mtllib, lone CR line
ends, faces that fail the load).  Compared: the flattened per-index stream (position, normal, texcoord bits, material of the
face) after load_model's vertex dedupe, and the material constants.

The golden digests (tests/golden/obj_fuzz_streams.json, tools/make_obj_fuzz_golden.py) were produced by that loader compiled
from the reference tree, so the comparison runs on any box; where oracle/_ref/tinyobj_dump exists the loader is also run live
on further seeds."""
import json
import os
import subprocess

import pytest

import obj_fuzz
from voxel_cone_tracing_b200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "obj_fuzz_streams.json")))["cases"]


def python_reader(path):
    try:
        return obj_fuzz.mesh_streams(S.load_obj(path))
    except S.ObjError:
        return None


def cpp_reader(path):
    if not os.path.exists(obj_fuzz.CPP_DUMP):
        pytest.skip("host/obj_dump not built (make)")
    out = path[:-4] + ".vctmesh"
    if subprocess.run([obj_fuzz.CPP_DUMP, path, out], capture_output=True).returncode:
        return None
    return obj_fuzz.mesh_streams(S.load_vctmesh(out))


READERS = {"python": python_reader, "cpp": cpp_reader}


@pytest.mark.parametrize("reader", sorted(READERS))
def test_reader_matches_reference_loader_digests(tmp_path, reader):
    bad = []
    for seed in range(len(GOLD)):
        got = obj_fuzz.digest(READERS[reader](obj_fuzz.write_case(str(tmp_path), seed)))
        if got != GOLD[str(seed)]:
            bad.append((seed, got, GOLD[str(seed)]))
    assert not bad, f"{len(bad)} of {len(GOLD)} files read differently from tinyobjloader: {bad[:5]}"
    assert sum(v == "no model" for v in GOLD.values()) < len(GOLD) // 10      # the generator mostly writes loadable files


@pytest.mark.parametrize("reader", sorted(READERS))
def test_reader_matches_live_reference_loader(tmp_path, reader):
    if not os.path.exists(obj_fuzz.TINYOBJ_DUMP):
        pytest.skip("oracle/_ref/tinyobj_dump not present (built from the reference tree)")
    for seed in range(len(GOLD)):        # the committed digests are what the loader says today
        if seed % 10 == 0:
            assert obj_fuzz.digest(obj_fuzz.tinyobj_streams(obj_fuzz.write_case(str(tmp_path), seed))) == GOLD[str(seed)], seed
    bad = []
    for seed in range(1000, 1120):       # seeds the digests do not cover, arrays compared directly
        path = obj_fuzz.write_case(str(tmp_path), seed)
        ref, got = obj_fuzz.tinyobj_streams(path), READERS[reader](path)
        if obj_fuzz.digest(ref) != obj_fuzz.digest(got):
            bad.append(seed)
    assert not bad, f"seeds {bad} read differently from tinyobjloader"


def test_float_parser_is_the_loaders_not_strtod():
    """Known answers of tryParseDouble (tiny_obj_loader.h:498-605) that strtod would read differently."""
    f = S._try_parse_double
    assert f(".5") is None and f("-.5") is None and f("abc") is None and f("+") is None and f("") is None
    assert f("1e") is None and f("1e+") is None            # an empty exponent fails the whole number
    assert f("5.") == 5.0 and f("1,5") == 1.0 and f("1.5f") == 1.5 and f("0x10") == 0.0 and f("1..2") == 1.0
    assert f("2e-3.5") == 2e-3 and f("1E2") == 100.0 and f("007") == 7.0 and f("-0") == 0.0 and str(f("-0")) == "-0.0"
    assert f("1e400") == float("inf") and f("1e-400") == 0.0
    # parseReal: a token that does not parse is the default 0, and the float is the double rounded once
    assert S._parse_real("v .5 2", 1)[0] == 0.0
    assert float(S._parse_real("0.1", 0)[0]) == float(__import__("numpy").float32(0.1))


def test_shape_is_lost_when_usemtl_directly_precedes_o(tmp_path):
    """The loader's `o` keeps the shape only if faces are pending: a material change right in front of it has already moved them."""
    (tmp_path / "m.mtl").write_text("newmtl a\nKd 1 0 0\nnewmtl b\nKd 0 1 0\n")
    (tmp_path / "lost.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nusemtl a\nf 1 2 3\nusemtl b\no second\nf 2 4 3\n")
    (tmp_path / "kept.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nusemtl a\nf 1 2 3\nusemtl b\ng second\nf 2 4 3\n")
    lost, kept = S.load_obj(str(tmp_path / "lost.obj")), S.load_obj(str(tmp_path / "kept.obj"))
    assert lost.ranges == [(0, 3, 1)] and kept.ranges == [(0, 3, 0), (3, 3, 1)]
    for name, mesh in (("lost.obj", lost), ("kept.obj", kept)):
        got = cpp_reader(str(tmp_path / name))
        assert obj_fuzz.digest(got) == obj_fuzz.digest(obj_fuzz.mesh_streams(mesh))
        if os.path.exists(obj_fuzz.TINYOBJ_DUMP):
            assert obj_fuzz.digest(obj_fuzz.tinyobj_streams(str(tmp_path / name))) == obj_fuzz.digest(got)


def test_load_failures(tmp_path):
    """A zero index fails the load (LoadObj returns false -> "Error loading file", INVALID_ID); so does an element that does not exist
    (undefined behaviour in the reference) and a file without triangles."""
    for i, body in enumerate(["v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 0\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n", "v 0 0 0\nv 1 0 0\nf 1 2\n", "# nothing\n"]):
        p = tmp_path / f"bad{i}.obj"
        p.write_text(body)
        with pytest.raises(S.ObjError):
            S.load_obj(str(p))
        assert cpp_reader(str(p)) is None
