"""Small scenes for the corners of the path (shared by the CPU test against the reference's GLSL and the GPU parity test): the 4-bit
count wrap, geometry that leaves the cube, degenerate triangles and NaN normals, more lights than MAX_POINT_LIGHTS with a
transmissive material, and a scene that produces nothing."""
import numpy as np

from voxel_cone_tracing_b200 import scene as S

EDGE_KINDS = ["stack", "outside", "degenerate", "lights", "tir", "mirror", "nolight", "empty"]


def _quad_mesh(z, half=0.2, copies=1):
    v = np.zeros(4, S.VERTEX)
    v["pos"] = [(-half, -half, z), (half, -half, z), (half, half, z), (-half, half, z)]
    v["norm"] = (0, 0, 1)
    return S.Mesh(v, np.array([0, 1, 2, 0, 2, 3] * copies, "<u4"), [(0, 6 * copies, -1)], np.zeros(0, S.MATERIAL))


def edge_scene(kind):
    b = S.SceneBuilder(1.0)
    m = S.default_material(); m["diffuse"][:3] = (0.3, 0.6, 0.9); m["emission"] = (0.1, 0.0, 0.2)
    if kind == "stack":            # 80 fragments per voxel: the 4-bit count wraps at 16 (voxelize.frag:80-93)
        b.add_mesh(_quad_mesh(0.1, copies=40), material_override=b.add_material(m))
        b.add_light((0.0, 0.0, 0.8))
    elif kind == "outside":        # geometry that leaves the cube: out-of-range image coordinates are dropped by imageAtomicCompSwap
        b.add_mesh(_quad_mesh(0.3, half=1.7), material_override=b.add_material(m))
        b.add_mesh(_quad_mesh(1.5, half=0.5), material_override=0)
        b.add_light((0.2, 0.1, 0.9), (1.0, 0.5, 0.25), 2.0)
    elif kind == "degenerate":     # zero-area and sliver triangles, a triangle exactly in a voxel plane, a normal of length zero
        v = np.zeros(9, S.VERTEX)
        v["pos"] = [(0, 0, 0), (0.5, 0.5, 0), (0.25, 0.25, 0),  (0.1, 0.1, 0.5), (0.9, 0.1, 0.5), (0.5, 0.100001, 0.5),  (-0.5, -0.5, 0.25), (0.5, -0.5, 0.25), (0.0, 0.5, 0.25)]
        v["norm"][:6] = (0, 0, 1)
        b.add_mesh(S.Mesh(v, np.arange(9, dtype="<u4"), [(0, 9, -1)], np.zeros(0, S.MATERIAL)), material_override=b.add_material(m))
        b.add_light((0.0, 0.0, 0.8))
    elif kind == "lights":         # twelve lights: the shader clamps to MAX_POINT_LIGHTS = 10; transmissive material (illum 7)
        m["illum"] = 7; m["dissolve"] = 0.35; m["transmittance"][:3] = (0.9, 0.4, 0.7); m["ior"] = 1.5; m["shininess"] = 40.0; m["specular"][:3] = (0.5, 0.5, 0.5)
        b.add_mesh(_quad_mesh(-0.1, half=0.6), mat_trs_tilt(), material_override=b.add_material(m))
        for i in range(12):
            b.add_light((0.7 * np.cos(i), 0.7 * np.sin(i), 0.5 + 0.03 * i), (1.0, 0.2 + 0.06 * i, 0.9 - 0.05 * i), 0.3)
    elif kind == "tir":            # ior < 1: refract() returns vec3(0) under total reflection, normalize() makes it NaN, the cone and the pixel with it
        m["illum"] = 4; m["dissolve"] = 0.5; m["transmittance"][:3] = (0.8, 0.9, 0.6); m["ior"] = 0.6; m["specular"][:3] = (0.4, 0.4, 0.4)
        b.add_mesh(_quad_mesh(0.0, half=0.7), S.mat_trs((0.0, 0.0, 0.0), 0.75, 1.0), material_override=b.add_material(m))
        b.add_mesh(_quad_mesh(-0.6, half=0.9), material_override=0)
        b.add_light((0.3, 0.4, 0.8))
    elif kind == "mirror":         # general model matrices: mirrored (negative determinant), non-uniformly scaled, sheared and rotated about x
        import os
        mesh = S.load_vctmesh(os.path.join(S.ASSET_DIR, "suzanne.vctmesh"))
        mid = b.add_material(m)
        c, s_ = np.cos(0.5), np.sin(0.5)
        rot_x = np.array([[1, 0, 0, 0], [0, c, -s_, 0], [0, s_, c, 0], [0, 0, 0, 1]], np.float64)
        shear = np.array([[-0.45, 0.12, 0, 0.2], [0, 0.3, 0.05, -0.1], [0.07, 0, 0.5, 0.05], [0, 0, 0, 1]], np.float64)   # det < 0
        model = (rot_x @ shear).astype(np.float32).T.reshape(16).copy()          # column-major like glm::mat4
        b.add_mesh(mesh, model, material_override=mid)
        m2 = S.default_material(); m2["diffuse"][:3] = (0.8, 0.7, 0.2); m2["specular"][:3] = (0.6, 0.6, 0.6); m2["shininess"] = 0.0
        b.add_mesh(_quad_mesh(-0.5, half=0.8), S.mat_trs((0.0, 0.0, 0.0), -0.3, 1.1), material_override=b.add_material(m2))
        b.add_light((0.4, 0.5, 0.9), (1.0, 0.9, 0.8), 1.5)
    elif kind == "nolight":        # point_light_count = 0: no shadow cones, direct term = emission
        b.add_mesh(_quad_mesh(0.0, half=0.5), S.mat_trs((0.0, 0.0, 0.0), 0.4, 1.0), material_override=b.add_material(m))
    elif kind == "empty":          # nothing inside the cube, nothing in front of the camera
        b.add_mesh(_quad_mesh(5.0, half=0.3), material_override=b.add_material(m))
        b.add_light((0.0, 0.0, 0.8))
    return b.build()


def mat_trs_tilt():
    return S.mat_trs((0.05, -0.1, 0.0), 0.6, 0.9)
