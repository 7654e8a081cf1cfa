"""The C++ host layer (include/vct/renderer.h, device.h, texture_3d.h -> libvct_host.so, vct_demo): the
reference's Renderer / Device / texture_3d API surface on top of the C ABI.

CPU part: the OBJ/MTL reader behind Renderer::load_model against the committed fixtures (and, where the
reference tree is present, against its OBJ files); the library refuses to render without a GPU.
GPU part: the headless demo (the reference's main() scene through the Renderer API) must produce the same
frame as the ctypes pipeline on the same inputs, and pass the frame gate against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from voxel_cone_tracing_b200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "voxel_cone_tracing_b200", "host")
DEMO = os.path.join(HOST, "vct_demo")
OBJ_DUMP = os.path.join(HOST, "obj_dump")
REF_ASSETS = "/root/reference/assets"


def test_host_binaries_built():
    for p in (DEMO, OBJ_DUMP, os.path.join(ROOT, "voxel_cone_tracing_b200", "libvct_host.so")):
        assert os.path.exists(p), f"{p} missing: run `make`"


@pytest.mark.parametrize("fixture", ["cornell_glossy.vctmesh", "suzanne.vctmesh"])
def test_cpp_reader_roundtrips_fixture(tmp_path, fixture):
    """load_model accepts the binary fixtures; reading and re-dumping one must be the identity"""
    src = os.path.join(ROOT, "assets", fixture)
    out = tmp_path / fixture
    subprocess.check_call([OBJ_DUMP, src, str(out)], stdout=subprocess.DEVNULL)
    assert out.read_bytes() == open(src, "rb").read()


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference assets not present (GPU box)")
@pytest.mark.parametrize("obj,fixture", [("CornellBox-Glossy.obj", "cornell_glossy.vctmesh"), ("suzanne.obj", "suzanne.vctmesh")])
def test_cpp_obj_reader_matches_python_reader_and_fixture(tmp_path, obj, fixture):
    """C++ OBJ/MTL reader == Python reader (itself pinned against tinyobjloader, test_scene_inputs.py) == committed fixture"""
    out = tmp_path / fixture
    subprocess.check_call([OBJ_DUMP, os.path.join(REF_ASSETS, obj), str(out)], stdout=subprocess.DEVNULL)
    got = S.load_vctmesh(str(out))
    exp = S.load_obj(os.path.join(REF_ASSETS, obj))
    assert np.array_equal(got.verts, exp.verts) and np.array_equal(got.indices, exp.indices)
    assert [tuple(r) for r in got.ranges] == [tuple(r) for r in exp.ranges]
    assert got.materials.tobytes() == exp.materials.tobytes()
    assert out.read_bytes() == open(os.path.join(ROOT, "assets", fixture), "rb").read()


def test_cpp_reader_reports_missing_file(tmp_path):
    r = subprocess.run([OBJ_DUMP, str(tmp_path / "nope.obj"), str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "Error loading file" in r.stderr      # renderer.cpp:417-421 message


def test_demo_refuses_to_run_without_gpu(gpu_available):
    if gpu_available:
        pytest.skip("GPU present")
    r = subprocess.run([DEMO, "--assets", os.path.join(ROOT, "assets")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


def _run_demo(tmp_path, *args):
    raw, dump = tmp_path / "frame.rgba", tmp_path / "scene.bin"
    out = subprocess.run([DEMO, "--assets", os.path.join(ROOT, "assets"), "--raw", str(raw), "--dump", str(dump), *args],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    mats = np.fromfile(str(dump), "<f4").reshape(3, 16)
    return np.fromfile(str(raw), "<u4"), mats, out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("suzanne,sampler", [(False, 0), (True, 0), (True, 1)])
def test_renderer_api_frame_equals_c_abi_pipeline_and_oracle(tmp_path, suzanne, sampler):
    from oracle import orc
    from voxel_cone_tracing_b200 import capi
    R, W, H = 64, 320, 240
    args = ["--res", str(R), "--size", f"{W}x{H}", "--sampler", str(sampler), "--frames", "3", "--theta", "0.4"]
    frame, mats, stdout = _run_demo(tmp_path, *(args + (["--suzanne"] if suzanne else [])))
    frame = frame.reshape(H, W)
    view, proj, dyn = mats
    assert "trace=" in stdout
    sc = S.cornell_scene(with_suzanne=suzanne)
    if suzanne:
        sc.draws["model"][-1] = dyn          # the matrix the C++ side built with vct::translate/rotate/scale (frame 2: theta + 0.1)
    # same inputs through the ctypes pipeline: bit-identical frame
    p = capi.Pipeline(sc, R, W, H)
    p.render_frame(view, proj, capi.default_params(sampler=sampler))
    assert np.array_equal(p.target.frame(), frame)
    p.close()
    # and inside the gate against the oracle
    ref = orc.render_frame(sc, view, proj, R, W, H)["frame"]
    d = np.abs(frame.view(np.uint8).astype(np.int32) - ref.view(np.uint8).astype(np.int32))
    assert d.max() <= 2
    # camera constants of the reference: P00, P11 at 4:3 and view column 3
    assert abs(proj[0] - 1.34444) < 1e-4 and abs(proj[5] - 1.79259) < 1e-4
    assert np.allclose(view[12:15], (0.0, -0.9, -3.0), atol=1e-6)


@pytest.mark.gpu
def test_renderer_api_voxel_debug_view(tmp_path):
    from oracle import orc
    R, W, H = 64, 256, 192
    frame, mats, _ = _run_demo(tmp_path, "--res", str(R), "--size", f"{W}x{H}", "--view-dir", "1", "--view-lod", "1.5")
    sc = S.cornell_scene()
    ref = orc.render_frame(sc, mats[0], mats[1], R, W, H, orc.default_params(view_voxel_dir=1, view_voxel_lod=1.5))["frame"]
    d = np.abs(frame.reshape(H, W).view(np.uint8).astype(np.int32) - ref.view(np.uint8).astype(np.int32))
    assert d.max() <= 2
