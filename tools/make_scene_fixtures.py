#!/usr/bin/env python
"""Regenerates assets/*.vctmesh from the reference's OBJ/MTL assets.

The reference tree (/root/reference) does not exist on the GPU box, so the benchmark scenes
are committed as small binary fixtures (format: voxel_cone_tracing_b200/scene.py, VCTMESH1).
They are INPUT DATA (vertex positions / normals / material constants of the public-domain
McGuire Cornell box and Blender's Suzanne), produced by our own OBJ reader; the reader is
cross-checked against the reference's vendored tinyobjloader in tests/test_scene_inputs.py.

    python tools/make_scene_fixtures.py [/root/reference/assets] [assets]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxel_cone_tracing_b200.scene import load_obj, save_vctmesh  # noqa: E402


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
    os.makedirs(dst, exist_ok=True)
    for obj, out in (("CornellBox-Glossy.obj", "cornell_glossy.vctmesh"), ("suzanne.obj", "suzanne.vctmesh")):
        m = load_obj(os.path.join(src, obj))
        save_vctmesh(m, os.path.join(dst, out))
        print(f"{obj}: {len(m.verts)} verts, {len(m.indices)//3} tris, {len(m.ranges)} ranges, {len(m.materials)} materials -> {out}")


if __name__ == "__main__":
    main()
