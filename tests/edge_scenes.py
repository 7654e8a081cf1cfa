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


# --------------------------------------------------------------------------- seeded random scenes
def fuzz_case(seed: int, big: bool = False):
    """A random small workload: triangle soups under random affine model matrices (some mirrored), random materials (opaque and
    transmissive, ior below and above 1, shininess 0 .. 1000, emission above 1), vertices partly outside the cube, a few zero normals,
    0 .. 12 lights inside and outside the cube, a random camera (often inside the geometry) and random phase toggles.
    -> (scene, R, levels, W, H, camera kwargs, trace-parameter kwargs)"""
    rng = np.random.default_rng(1000 + seed)
    cube = float(rng.choice([0.75, 1.0, 2.0, 3.0]))
    b = S.SceneBuilder(cube)
    n_mats = int(rng.integers(1, 4))
    for _ in range(n_mats):
        m = S.default_material()
        m["diffuse"][:3] = rng.random(3); m["specular"][:3] = rng.random(3) * rng.choice([0.0, 1.0]); m["transmittance"][:3] = rng.random(3)
        m["emission"] = rng.random(3) * rng.choice([0.0, 0.3, 1.5])
        m["shininess"] = float(rng.choice([0.0, 1.0, 10.0, 96.0, 1000.0])); m["ior"] = float(rng.choice([0.6, 1.0, 1.45, 5.0]))
        m["dissolve"] = float(rng.choice([0.0, 0.05, 0.1, 0.5, 1.0])); m["illum"] = int(rng.choice([0, 2, 2, 4, 6, 7, 9]))
        b.add_material(m)
    for _ in range(int(rng.integers(1, 4))):
        n_tri = int(rng.integers(8, 160))
        centre = (rng.random((n_tri, 1, 3)) * 2.4 - 1.2) * cube
        size = rng.choice([0.6, 1.5, 2.5] if big else [0.05, 0.2, 0.6], (n_tri, 1, 1)) * cube   # big: dozens of fragments per voxel (count wrap)
        pos = (centre + (rng.random((n_tri, 3, 3)) - 0.5) * size).reshape(-1, 3)
        v = np.zeros(3 * n_tri, S.VERTEX)
        v["pos"] = pos
        nrm = rng.standard_normal((3 * n_tri, 3))
        nrm[rng.random(3 * n_tri) < 0.02] = 0.0                      # a few normals of length zero (NaN after normalize)
        v["norm"] = nrm
        v["uv"] = rng.random((3 * n_tri, 2))
        a = rng.standard_normal((3, 3)) * 0.35 + np.eye(3) * rng.choice([-1.0, 1.0]) * 0.8
        model = np.eye(4); model[:3, :3] = a; model[:3, 3] = (rng.random(3) - 0.5) * 0.3 * cube
        b.add_mesh(S.Mesh(v, np.arange(3 * n_tri, dtype="<u4"), [(0, 3 * n_tri, -1)], np.zeros(0, S.MATERIAL)),
                   model.astype(np.float32).T.reshape(16).copy(), material_override=int(rng.integers(0, n_mats)))
    for _ in range(int(rng.choice([0, 1, 1, 2, 3, 12]))):
        b.add_light((rng.random(3) * 2.6 - 1.3) * cube, rng.random(3), float(rng.choice([0.5, 1.0, 2.0])))
    R = int(rng.choice([32, 64]))
    eye = (rng.random(3) * 2.0 - 1.0) * cube * 1.2
    d = -eye / max(float(np.linalg.norm(eye)), 1e-3)                  # towards the middle of the cube, +- 20 degrees
    cam = dict(eye=tuple(eye.tolist()), pitch=float(np.degrees(np.arcsin(np.clip(d[1], -1, 1))) + rng.uniform(-20, 20)),
               yaw=float(np.degrees(np.arctan2(d[2], d[0])) + rng.uniform(-20, 20)))
    prm = dict(enable_direct=int(rng.random() < 0.85), enable_diffuse=int(rng.random() < 0.85), enable_specular=int(rng.random() < 0.85),
               enable_shadow=int(rng.random() < 0.8))
    if rng.random() < 0.15:
        prm.update(view_voxel_dir=int(rng.integers(0, 6)), view_voxel_lod=float(rng.choice([0.0, 0.5, 2.25, 9.0])))
    return b.build(), R, (6 if R == 32 else 7), int(rng.choice([64, 96, 131])), int(rng.choice([48, 64, 77])), cam, prm


FUZZ_SEEDS = list(range(24))
