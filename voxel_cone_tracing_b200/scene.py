"""Host-side scene description: the POD records that cross the C ABI, the binary mesh
fixture format, an OBJ/MTL reader, the glm-compatible camera maths and the benchmark scenes.

Mirrors the reference's host types (src/renderer.cpp:24-35 vert_data_t, src/renderer.h:57-120
draw_obj_t / model_t / point_light_t / material_data_t) and scene constants
(src/main.cpp:85-121, src/camera.h:12-58).  No compute happens here.
"""
from __future__ import annotations

import math
import os
import struct
from dataclasses import dataclass, field

import numpy as np

VERTEX = np.dtype([("pos", "<f4", 3), ("norm", "<f4", 3), ("uv", "<f4", 2)])
DRAW = np.dtype([("first_index", "<u4"), ("index_count", "<u4"), ("vertex_base", "<u4"),
                 ("material", "<u4"), ("model", "<f4", 16)])
LIGHT = np.dtype([("position", "<f4", 3), ("color", "<f4", 3), ("intensity", "<f4")])
MATERIAL = np.dtype([("ambient", "<f4", 4), ("diffuse", "<f4", 4), ("specular", "<f4", 4),
                     ("transmittance", "<f4", 4), ("emission", "<f4", 3), ("shininess", "<f4"),
                     ("ior", "<f4"), ("dissolve", "<f4"), ("illum", "<i4"), ("roughness", "<f4"),
                     ("metallic", "<f4"), ("sheen", "<f4"), ("clearcoat_thickness", "<f4"),
                     ("clearcoat_roughness", "<f4"), ("anisotropy", "<f4"),
                     ("anisotropy_rotation", "<f4"), ("pad", "<f4", 2)])
assert VERTEX.itemsize == 32 and DRAW.itemsize == 80 and LIGHT.itemsize == 28 and MATERIAL.itemsize == 128

ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
MESH_MAGIC = b"VCTMESH1"


# ----------------------------------------------------------------------------- meshes
@dataclass
class Mesh:
    """One loaded model: deduplicated vertices, indices, per-material index ranges."""
    verts: np.ndarray            # VERTEX[n]
    indices: np.ndarray          # u32[m]
    ranges: list                 # [(first_index, index_count, local_material or -1)]
    materials: np.ndarray        # MATERIAL[k]  (from the .mtl, in file order)
    material_names: list = field(default_factory=list)


def save_vctmesh(mesh: Mesh, path: str) -> None:
    with open(path, "wb") as f:
        f.write(MESH_MAGIC)
        f.write(struct.pack("<4I", len(mesh.verts), len(mesh.indices), len(mesh.ranges), len(mesh.materials)))
        f.write(np.ascontiguousarray(mesh.verts).tobytes())
        f.write(np.ascontiguousarray(mesh.indices, dtype="<u4").tobytes())
        for (a, b, m) in mesh.ranges:
            f.write(struct.pack("<IIi", a, b, m))
        f.write(np.ascontiguousarray(mesh.materials).tobytes())


def load_vctmesh(path: str) -> Mesh:
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:8] != MESH_MAGIC:
        raise ValueError(f"{path}: not a VCTMESH1 file")
    nv, ni, nr, nm = struct.unpack_from("<4I", blob, 8)
    o = 24
    verts = np.frombuffer(blob, VERTEX, nv, o).copy(); o += nv * 32
    indices = np.frombuffer(blob, "<u4", ni, o).copy(); o += ni * 4
    ranges = [struct.unpack_from("<IIi", blob, o + 12 * i) for i in range(nr)]; o += nr * 12
    mats = np.frombuffer(blob, MATERIAL, nm, o).copy()
    return Mesh(verts, indices, ranges, mats)


def default_material() -> np.ndarray:
    """tinyobjloader InitMaterial defaults (thirdparty/tinyobjloader/tiny_obj_loader.h:936-957)."""
    m = np.zeros((), MATERIAL)
    m["dissolve"] = 1.0
    m["shininess"] = 1.0
    m["ior"] = 1.0
    return m


def parse_mtl(path: str):
    """MTL subset used by the reference's create_material (src/renderer.cpp:49-81)."""
    mats, names = [], []
    cur = None
    has_d = False
    with open(path, "r", errors="replace") as f:
        for raw in f:
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            tok = line.split()
            key, args = tok[0], tok[1:]
            if key == "newmtl":
                if cur is not None:
                    mats.append(cur)
                cur = default_material()
                names.append(args[0] if args else "")
                has_d = False
                continue
            if cur is None:
                continue
            f3 = lambda: [float(a) for a in (args + ["0", "0", "0"])[:3]]
            if key == "Ka": cur["ambient"][:3] = f3()
            elif key == "Kd": cur["diffuse"][:3] = f3()
            elif key == "Ks": cur["specular"][:3] = f3()
            elif key in ("Kt", "Tf"): cur["transmittance"][:3] = f3()
            elif key == "Ke": cur["emission"][:3] = f3()
            elif key == "Ni": cur["ior"] = float(args[0])
            elif key == "Ns": cur["shininess"] = float(args[0])
            elif key == "illum": cur["illum"] = int(float(args[0]))
            elif key == "d":
                cur["dissolve"] = float(args[0]); has_d = True
            elif key == "Tr":
                if not has_d:                      # `d` wins over `Tr` (tiny_obj_loader.h:1203-1222)
                    cur["dissolve"] = 1.0 - float(args[0])
            elif key == "Pr": cur["roughness"] = float(args[0])
            elif key == "Pm": cur["metallic"] = float(args[0])
            elif key == "Ps": cur["sheen"] = float(args[0])
            elif key == "Pc": cur["clearcoat_thickness"] = float(args[0])
            elif key == "Pcr": cur["clearcoat_roughness"] = float(args[0])
            elif key == "aniso": cur["anisotropy"] = float(args[0])
            elif key == "anisor": cur["anisotropy_rotation"] = float(args[0])
    if cur is not None:
        mats.append(cur)
    arr = np.array(mats, MATERIAL) if mats else np.zeros(0, MATERIAL)
    return arr, names


def load_obj(path: str) -> Mesh:
    """OBJ -> deduplicated vertex/index buffers + per-material ranges, following
    Renderer::load_model (src/renderer.cpp:407-603): faces fan-triangulated, one vertex per
    distinct (pos, normal, uv) triple in first-use order, a new range whenever the material
    changes inside a shape or a new group/object starts."""
    pos, nrm, tex = [], [], []
    mats = np.zeros(0, MATERIAL)
    names: list = []
    cur_mat = -1
    verts, index, uniq = [], [], {}
    ranges = []
    start = 0
    range_mat = None
    base = os.path.dirname(path)

    def close_range():
        nonlocal start, range_mat
        if len(index) > start:
            ranges.append((start, len(index) - start, range_mat if range_mat is not None else -1))
        start = len(index)
        range_mat = None

    with open(path, "r", errors="replace") as f:
        for raw in f:
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            tok = line.split()
            key, args = tok[0], tok[1:]
            if key == "v": pos.append(tuple(np.float32(a) for a in args[:3]))
            elif key == "vn": nrm.append(tuple(np.float32(a) for a in args[:3]))
            elif key == "vt": tex.append(tuple(np.float32(a) for a in (args + ["0"])[:2]))
            elif key == "mtllib":
                mats, names = parse_mtl(os.path.join(base, args[0]))
            elif key == "usemtl":
                cur_mat = names.index(args[0]) if args and args[0] in names else -1
            elif key in ("g", "o"):
                close_range()
            elif key == "f":
                corners = []
                for a in args:
                    parts = (a.split("/") + ["", ""])[:3]
                    vi = int(parts[0]); vi = vi - 1 if vi > 0 else len(pos) + vi
                    ti = None
                    if parts[1]:
                        ti = int(parts[1]); ti = ti - 1 if ti > 0 else len(tex) + ti
                    ni = None
                    if parts[2]:
                        ni = int(parts[2]); ni = ni - 1 if ni > 0 else len(nrm) + ni
                    corners.append((vi, ti, ni))
                if range_mat is not None and range_mat != cur_mat:
                    close_range()
                range_mat = cur_mat
                for k in range(1, len(corners) - 1):
                    for c in (corners[0], corners[k], corners[k + 1]):
                        p = pos[c[0]]
                        t = tex[c[1]] if c[1] is not None else (np.float32(0), np.float32(0))
                        n = nrm[c[2]] if c[2] is not None else (np.float32(0),) * 3
                        keyv = (p, n, t)
                        j = uniq.get(keyv)
                        if j is None:
                            j = len(verts); uniq[keyv] = j; verts.append(keyv)
                        index.append(j)
    close_range()
    varr = np.zeros(len(verts), VERTEX)
    for i, (p, n, t) in enumerate(verts):
        varr[i]["pos"] = p; varr[i]["norm"] = n; varr[i]["uv"] = t
    return Mesh(varr, np.array(index, "<u4"), ranges, mats, names)


# ----------------------------------------------------------------------------- camera (glm 0.9.9 semantics)
def perspective(fovy: float, aspect: float, zn: float, zf: float) -> np.ndarray:
    """glm::perspective RH, depth -1..1; fovy in RADIANS (the reference passes 45.0f: camera.h:23, main.cpp:108)."""
    t = np.float32(math.tan(np.float32(fovy) / np.float32(2)))
    m = np.zeros(16, np.float32)
    m[0] = np.float32(1) / (np.float32(aspect) * t)
    m[5] = np.float32(1) / t
    m[10] = -(np.float32(zf) + np.float32(zn)) / (np.float32(zf) - np.float32(zn))
    m[11] = -1.0
    m[14] = -(np.float32(2) * np.float32(zf) * np.float32(zn)) / (np.float32(zf) - np.float32(zn))
    return m


def _norm(v):
    v = np.asarray(v, np.float32)
    return v / np.float32(np.sqrt(np.float32(np.dot(v, v))))


def look_at(eye, center, up) -> np.ndarray:
    eye = np.asarray(eye, np.float32)
    f = _norm(np.asarray(center, np.float32) - eye)
    s = _norm(np.cross(f, np.asarray(up, np.float32)).astype(np.float32))
    u = np.cross(s, f).astype(np.float32)
    m = np.zeros(16, np.float32)
    m[0], m[4], m[8] = s
    m[1], m[5], m[9] = u
    m[2], m[6], m[10] = -f
    m[12] = -np.dot(s, eye); m[13] = -np.dot(u, eye); m[14] = np.dot(f, eye); m[15] = 1.0
    return m


def camera_front(pitch_deg: float, yaw_deg: float) -> np.ndarray:
    """Camera::calc_front (src/camera.h:25-37)."""
    p, y = np.float32(math.radians(pitch_deg)), np.float32(math.radians(yaw_deg))
    f = np.array([np.cos(p) * np.cos(y), np.sin(p), np.cos(p) * np.sin(y)], np.float32)
    return _norm(f)


def reference_camera(aspect: float, eye=(0.0, 0.9, 3.0), pitch=0.0, yaw=-90.0):
    """The reference's camera (src/main.cpp:107-109): returns (view, projection) column-major."""
    eye = np.asarray(eye, np.float32)
    front = camera_front(pitch, yaw)
    view = look_at(eye, eye + front, (0.0, 1.0, 0.0))
    proj = perspective(45.0, aspect, 0.1, 100.0)
    return view, proj


def mat_identity(): return np.eye(4, dtype=np.float32).T.reshape(16).copy()


def mat_trs(t, ry: float, s: float) -> np.ndarray:
    """translate(t) * rotate(ry about +y) * scale(s), column-major (src/main.cpp:369-372)."""
    c, sn = np.float32(math.cos(ry)), np.float32(math.sin(ry))
    m = np.zeros((4, 4), np.float32)  # m[col][row]
    m[0] = [c * s, 0, -sn * s, 0]
    m[1] = [0, s, 0, 0]
    m[2] = [sn * s, 0, c * s, 0]
    m[3] = [t[0], t[1], t[2], 1]
    return m.reshape(16).copy()


# ----------------------------------------------------------------------------- scenes
@dataclass
class Scene:
    verts: np.ndarray
    indices: np.ndarray
    draws: np.ndarray
    materials: np.ndarray
    lights: np.ndarray
    cube_size: float = 3.0

    @property
    def n_triangles(self) -> int:
        return int(self.draws["index_count"].sum() // 3)


class SceneBuilder:
    """Accumulates models the way Renderer does: material 0 is the default material
    (src/renderer.cpp:83-98,115-128), each load appends its materials after it."""

    def __init__(self, cube_size: float = 3.0):
        d = default_material()
        d["ambient"][:3] = (1, 0, 1); d["diffuse"][:3] = (1, 0, 1)   # fill_default_mat_data
        d["shininess"] = 1; d["ior"] = 1; d["dissolve"] = 1; d["illum"] = 0
        self.materials = [d]
        self.verts, self.indices, self.draws, self.lights = [], [], [], []
        self.nv = 0; self.ni = 0
        self.cube_size = cube_size

    def add_material(self, m) -> int:
        self.materials.append(np.array(m, MATERIAL)); return len(self.materials) - 1

    def add_mesh(self, mesh: Mesh, model=None, material_override=None) -> None:
        model = mat_identity() if model is None else np.asarray(model, np.float32)
        mbase = len(self.materials)
        for m in mesh.materials:
            self.materials.append(m)
        for (a, b, lm) in mesh.ranges:
            d = np.zeros((), DRAW)
            d["first_index"] = self.ni + a; d["index_count"] = b; d["vertex_base"] = self.nv
            d["material"] = material_override if material_override is not None else (mbase + lm if lm >= 0 else 0)
            d["model"] = model
            self.draws.append(d)
        self.verts.append(mesh.verts); self.indices.append(mesh.indices)
        self.nv += len(mesh.verts); self.ni += len(mesh.indices)

    def add_light(self, position, color=(1, 1, 1), intensity=1.0) -> None:
        l = np.zeros((), LIGHT); l["position"] = position; l["color"] = color; l["intensity"] = intensity
        self.lights.append(l)

    def build(self) -> Scene:
        return Scene(np.concatenate(self.verts).astype(VERTEX), np.concatenate(self.indices).astype("<u4"),
                     np.array(self.draws, DRAW), np.array(self.materials, MATERIAL),
                     np.array(self.lights, LIGHT) if self.lights else np.zeros(0, LIGHT), self.cube_size)


def suzanne_material() -> np.ndarray:
    """dynamic_object_material (src/main.cpp:88-97)."""
    m = np.zeros((), MATERIAL)
    m["ambient"] = 1.0; m["diffuse"] = 0.0; m["specular"] = 1.0; m["transmittance"] = 1.0
    m["emission"] = (0.0, 0.0, 0.25); m["shininess"] = 1000; m["ior"] = 5; m["dissolve"] = 0.1; m["illum"] = 4
    return m


def cornell_scene(with_suzanne: bool = False, theta: float = 0.0, asset_dir: str = ASSET_DIR) -> Scene:
    """The reference scene (src/main.cpp:85-121): Cornell box (identity), optionally Suzanne at
    T(0,1.1,-0.5)*Ry(theta)*S(0.3) with the refractive material, one white light at (0,1.4,0), cube_size 3."""
    b = SceneBuilder(3.0)
    b.add_mesh(load_vctmesh(os.path.join(asset_dir, "cornell_glossy.vctmesh")))
    if with_suzanne:
        mid = None
        mesh = load_vctmesh(os.path.join(asset_dir, "suzanne.vctmesh"))
        # add_material happens after both loads in the reference; ids only need to be consistent here
        b.add_mesh(mesh, mat_trs((0.0, 1.1, -0.5), theta, 0.3), material_override=0)
        mid = b.add_material(suzanne_material())
        for d in b.draws[-len(mesh.ranges):]:
            d["material"] = mid
    b.add_light((0.0, 1.4, 0.0), (1.0, 1.0, 1.0), 1.0)
    return b.build()


# ---- synthetic scenes (BASELINE configs 4/5): PCG32-seeded closed box + icospheres ----
class PCG32:
    def __init__(self, seed: int, seq: int = 54):
        self.state = 0; self.inc = ((seq << 1) | 1) & 0xFFFFFFFFFFFFFFFF
        self.next(); self.state = (self.state + seed) & 0xFFFFFFFFFFFFFFFF; self.next()

    def next(self) -> int:
        old = self.state
        self.state = (old * 6364136223846793005 + self.inc) & 0xFFFFFFFFFFFFFFFF
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF

    def uniform(self, lo=0.0, hi=1.0) -> float:
        return lo + (hi - lo) * (self.next() / 4294967296.0)


def _icosphere(subdiv: int):
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    v = [tuple(np.array(p) / np.linalg.norm(p)) for p in v]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                p = (np.array(v[a]) + np.array(v[b])) / 2.0
                v.append(tuple(p / np.linalg.norm(p))); cache[k] = len(v) - 1
            return cache[k]
        for (a, b, c) in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v, np.float32), np.array(f, np.uint32)


def synthetic_scene(n_triangles: int, seed: int, cube_size: float = 3.0, subdiv: int = 4) -> Scene:
    """Closed box (12 triangles, Cornell wall materials) + K icospheres (5120 triangles each at
    subdiv 4) at PCG32-seeded positions/radii inside [-0.9,0.9]^3 * cube_size."""
    rng = PCG32(seed)
    box_mats = load_vctmesh(os.path.join(ASSET_DIR, "cornell_glossy.vctmesh")).materials
    b = SceneBuilder(cube_size)
    for m in box_mats:
        b.materials.append(m)
    e = 0.95 * cube_size
    corners = np.array([[-e, -e, -e], [e, -e, -e], [e, e, -e], [-e, e, -e], [-e, -e, e], [e, -e, e], [e, e, e], [-e, e, e]], np.float32)
    faces = [((0, 1, 2, 3), (0, 0, 1), 5), ((5, 4, 7, 6), (0, 0, -1), 5), ((4, 0, 3, 7), (1, 0, 0), 7), ((1, 5, 6, 2), (-1, 0, 0), 6),
             ((4, 5, 1, 0), (0, 1, 0), 3), ((3, 2, 6, 7), (0, -1, 0), 4)]
    bv, bi, ranges = [], [], []
    for (q, n, mat) in faces:
        s = len(bv)
        for k in q:
            vv = np.zeros((), VERTEX); vv["pos"] = corners[k]; vv["norm"] = n; bv.append(vv)
        ranges.append((len(bi), 6, mat)); bi += [s, s + 1, s + 2, s, s + 2, s + 3]
    mesh = Mesh(np.array(bv, VERTEX), np.array(bi, "<u4"), [(a, c, -1) for (a, c, _) in ranges], np.zeros(0, MATERIAL))
    for (a, c, mat) in ranges:
        d = np.zeros((), DRAW); d["first_index"] = a; d["index_count"] = c; d["material"] = mat; d["model"] = mat_identity(); b.draws.append(d)
    b.verts.append(mesh.verts); b.indices.append(mesh.indices); b.nv = len(mesh.verts); b.ni = len(mesh.indices)
    sv, sf = _icosphere(subdiv)
    sphere = np.zeros(len(sv), VERTEX); sphere["pos"] = sv; sphere["norm"] = sv
    sidx = sf.reshape(-1).astype("<u4")
    k = max(0, (n_triangles - 12) // len(sf))
    b.verts.append(sphere); b.indices.append(sidx)
    for _ in range(k):
        r = rng.uniform(0.02, 0.12) * cube_size
        c = [rng.uniform(-0.9, 0.9) * cube_size * 0.9 for _ in range(3)]
        d = np.zeros((), DRAW)
        d["first_index"] = b.ni; d["index_count"] = len(sidx); d["vertex_base"] = b.nv
        d["material"] = 1 + (rng.next() % 8); d["model"] = mat_trs(c, rng.uniform(0, 6.2831853), r)
        b.draws.append(d)
    b.nv += len(sphere); b.ni += len(sidx)
    b.add_light((0.0, 0.8 * cube_size, 0.0), (1.0, 1.0, 1.0), 1.0)
    return b.build()
