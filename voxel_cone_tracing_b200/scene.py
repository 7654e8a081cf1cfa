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


# --- tinyobjloader v1.1.0's token grammar (thirdparty/tinyobjloader/tiny_obj_loader.h), restated -------------------------------
# The reference's models are whatever that loader makes of the file (src/renderer.cpp:417), including what it makes of
# malformed numbers and of statements in unusual order; tests/test_obj_reader_fuzz.py holds this reader and the C++ one
# (host/obj_loader.cpp) against the loader itself on random files.
_POW_LUT = (1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001)


def _isdigit(c: str) -> bool:
    return "0" <= c <= "9"


def _try_parse_double(s: str):
    """tryParseDouble (tiny_obj_loader.h:498-605): [sign] digits ['.' digits] [(e|E) [sign] digits], greedy, value of
    the conforming prefix; None where the loader reports failure (no leading digit, empty exponent).  The arithmetic is
    the loader's own (decimal digits added through a table of powers, 10^e as ldexp(m * 5^e, e)), not strtod's."""
    n = len(s)
    if n == 0:
        return None
    i, sign = 0, 1.0
    if s[0] in "+-":
        sign = -1.0 if s[0] == "-" else 1.0
        i = 1
    elif not _isdigit(s[0]):
        return None
    mant, read = 0.0, 0
    while i < n and _isdigit(s[i]):
        mant = mant * 10.0 + (ord(s[i]) - 48)
        i += 1
        read += 1
    if read == 0:
        return None
    exponent = 0
    if i < n and s[i] == ".":
        i += 1
        read = 1
        while i < n and _isdigit(s[i]):
            mant += (ord(s[i]) - 48) * (_POW_LUT[read] if read < 8 else math.pow(10.0, -read))
            read += 1
            i += 1
    if i < n and s[i] in "eE":
        i += 1
        esign = 1
        if i < n and s[i] in "+-":
            esign = -1 if s[i] == "-" else 1
            i += 1
        elif not (i < n and _isdigit(s[i])):
            return None
        read = 0
        while i < n and _isdigit(s[i]):
            exponent = exponent * 10 + (ord(s[i]) - 48)
            i += 1
            read += 1
        exponent *= esign
        if read == 0:
            return None
    if exponent:
        try:
            mant = math.ldexp(mant * math.pow(5.0, exponent), exponent)
        except OverflowError:
            mant = math.inf
    return sign * mant


def _skip_ws(line: str, pos: int) -> int:
    while pos < len(line) and line[pos] in " \t":
        pos += 1
    return pos


def _parse_real(line: str, pos: int, default: float = 0.0):
    """parseReal (:612-621): next blank-separated token as a float (double rounded once), `default` where it does not parse."""
    pos = _skip_ws(line, pos)
    end = pos
    while end < len(line) and line[end] not in " \t\r":
        end += 1
    val = _try_parse_double(line[pos:end])
    with np.errstate(over="ignore"):
        return np.float32(default if val is None else val), end


def _reals(line: str, pos: int, n: int):
    out = []
    for _ in range(n):
        v, pos = _parse_real(line, pos)
        out.append(v)
    return tuple(out)


def _atoi(line: str, pos: int) -> int:
    n = len(line)
    while pos < n and line[pos] in " \t\n\v\f\r":
        pos += 1
    sign = 1
    if pos < n and line[pos] in "+-":
        sign = -1 if line[pos] == "-" else 1
        pos += 1
    v = 0
    while pos < n and _isdigit(line[pos]):
        v = v * 10 + (ord(line[pos]) - 48)
        pos += 1
    return sign * v


def _split_lines(blob: bytes):
    """safeGetline: \n, \r\n and a lone \r all end a line."""
    return blob.decode("latin-1").replace("\r\n", "\n").replace("\r", "\n").split("\n")


def _is_stmt(tok: str, key: str) -> bool:
    return tok.startswith(key) and len(tok) > len(key) and tok[len(key)] in " \t"


def parse_mtl(path: str):
    """LoadMtl (tiny_obj_loader.h:1049-1431) for the constants create_material forwards (src/renderer.cpp:49-81):
    (materials, names).  A material's name is the rest of its `newmtl` line; statements in front of the first `newmtl`
    belong to a nameless material that is dropped, a file without any `newmtl` yields that one nameless material."""
    mats, names = [], []
    cur, name = default_material(), ""
    has_d = False
    with open(path, "rb") as f:
        lines = _split_lines(f.read())
    for raw in lines:
        tok = raw.rstrip(" \t").lstrip(" \t")
        if not tok or tok[0] == "#":
            continue
        if _is_stmt(tok, "newmtl"):
            if name:
                mats.append(cur); names.append(name)
            cur, name, has_d = default_material(), tok[7:], False
            continue
        three = None
        for key, fld in (("Ka", "ambient"), ("Kd", "diffuse"), ("Ks", "specular"), ("Kt", "transmittance"), ("Tf", "transmittance"), ("Ke", "emission")):
            if _is_stmt(tok, key):
                three = fld
                break
        if three:
            cur[three][:3] = _reals(tok, 2, 3)
            continue
        one = None
        for key, fld in (("Ni", "ior"), ("Ns", "shininess"), ("Pr", "roughness"), ("Pm", "metallic"), ("Ps", "sheen"), ("Pc", "clearcoat_thickness"),
                         ("Pcr", "clearcoat_roughness"), ("aniso", "anisotropy"), ("anisor", "anisotropy_rotation")):
            if _is_stmt(tok, key):
                one = (fld, len(key))
                break
        if one:
            cur[one[0]] = _parse_real(tok, one[1])[0]
        elif _is_stmt(tok, "illum"):
            cur["illum"] = _atoi(tok, 6)                         # parseInt = atoi
        elif _is_stmt(tok, "d"):
            cur["dissolve"] = _parse_real(tok, 1)[0]
            has_d = True
        elif _is_stmt(tok, "Tr"):
            if not has_d:                          # `d` wins over `Tr` (:1203-1222); the subtraction is in float
                cur["dissolve"] = np.float32(1.0) - _parse_real(tok, 2)[0]
    mats.append(cur); names.append(name)           # the last material is flushed whatever its name (:1426-1428)
    return np.array(mats, MATERIAL), names


class ObjError(ValueError):
    """tinyobj::LoadObj returned false (Renderer::load_model prints "Error loading file" and returns INVALID_ID)."""


def load_obj(path: str) -> Mesh:
    """OBJ -> deduplicated vertex/index buffers + per-material ranges.

    Parsing restates tinyobj::LoadObj with triangulate = true (tiny_obj_loader.h:1525-1790) as a state machine of its own:
    faces collect in a pending group; `usemtl` with a different material moves the group into the current shape (tagged
    with the material that was current), `g` and `o` do the same and then close the shape -- `g` keeps it if it has
    triangles, `o` keeps it only if the pending group was not empty (so `usemtl` directly in front of an `o` loses the
    shape: the loader's behaviour, reproduced) -- and the end of the file keeps whatever is left.  Polygons become fans.
    Flattening follows Renderer::load_model (src/renderer.cpp:407-557): one vertex per distinct (pos, normal, uv) triple
    in first-use order, one range per shape and per run of one material inside it (material -1 = none)."""
    pos, nrm, tex = [], [], []
    mats_list, names = [], []
    mat_map = {}
    material = -1
    group, shape, shapes = [], [], []
    base = path[: max(path.rfind("/"), path.rfind("\\")) + 1]     # renderer.cpp:413-414

    def export() -> bool:
        if not group:
            return False
        for face in group:
            for k in range(2, len(face)):
                shape.append((face[0], face[k - 1], face[k], material))
        return True

    with open(path, "rb") as f:
        lines = _split_lines(f.read())
    for raw in lines:
        tok = raw.lstrip(" \t")
        if not tok or tok[0] == "#":
            continue
        if _is_stmt(tok, "v"):
            pos.append(_reals(tok, 2, 3))
        elif _is_stmt(tok, "vn"):
            nrm.append(_reals(tok, 3, 3))
        elif _is_stmt(tok, "vt"):
            tex.append(_reals(tok, 3, 2))
        elif _is_stmt(tok, "f"):
            p, n, face = _skip_ws(tok, 2), len(tok), []

            def fix(idx: int, count: int):
                if idx == 0:
                    raise ObjError(f"{path}: zero index in `f` line")
                return idx - 1 if idx > 0 else count + idx

            def skip_index(q: int) -> int:
                while q < n and tok[q] not in "/ \t\r":
                    q += 1
                return q

            while p < n:                                           # parseTriple (:745-799): i, i/j, i//k, i/j/k
                vi, ti, ni = fix(_atoi(tok, p), len(pos)), None, None
                p = skip_index(p)
                if p < n and tok[p] == "/":
                    p += 1
                    if p < n and tok[p] == "/":
                        p += 1
                        ni = fix(_atoi(tok, p), len(nrm))
                        p = skip_index(p)
                    else:
                        ti = fix(_atoi(tok, p), len(tex))
                        p = skip_index(p)
                        if p < n and tok[p] == "/":
                            p += 1
                            ni = fix(_atoi(tok, p), len(nrm))
                            p = skip_index(p)
                # a relative normal / texcoord index that points in front of the array counts as "none" (load_model tests >= 0,
                # renderer.cpp:489,499); anything else out of range is undefined behaviour in the reference: refused here
                ti = None if ti is not None and ti < 0 else ti
                ni = None if ni is not None and ni < 0 else ni
                if not (0 <= vi < len(pos)) or (ti is not None and ti >= len(tex)) or (ni is not None and ni >= len(nrm)):
                    raise ObjError(f"{path}: `f` line refers to an element that does not exist")
                face.append((vi, ti, ni))
                p = _skip_ws(tok, p)
            group.append(face)
        elif _is_stmt(tok, "usemtl"):
            new = mat_map.get(tok[7:], -1)
            if new != material:
                export()
                group = []
                material = new
        elif _is_stmt(tok, "mtllib"):
            for name in tok[7:].split(" "):
                try:
                    m, nm = parse_mtl(base + name)
                except OSError:
                    continue                                       # "WARN: Material file not found", the load goes on
                for mm, nn in zip(m, nm):
                    mat_map.setdefault(nn, len(mats_list))
                    mats_list.append(mm); names.append(nn)
                break
        elif _is_stmt(tok, "g"):
            export()
            if shape:
                shapes.append(shape)
            shape, group = [], []
        elif _is_stmt(tok, "o"):
            if export():
                shapes.append(shape)
            shape, group = [], []
    if export() or shape:
        shapes.append(shape)

    verts, index, uniq, ranges = [], [], {}, []
    zero2, zero3 = (np.float32(0),) * 2, (np.float32(0),) * 3
    for sh in shapes:
        start, run_mat = len(index), None
        for (c0, c1, c2, m) in sh:
            if run_mat is not None and m != run_mat:
                ranges.append((start, len(index) - start, run_mat))
                start = len(index)
            run_mat = m
            for (vi, ti, ni) in (c0, c1, c2):
                keyv = (pos[vi], nrm[ni] if ni is not None else zero3, tex[ti] if ti is not None else zero2)
                j = uniq.get(keyv)                                 # -0.0 == 0.0 here as in vert_data_t::operator== (renderer.cpp:31-34); first-seen bits kept
                if j is None:
                    j = len(verts); uniq[keyv] = j; verts.append(keyv)
                index.append(j)
        if len(index) > start:
            ranges.append((start, len(index) - start, run_mat))
    if not index:
        raise ObjError(f"{path}: no triangles")
    varr = np.zeros(len(verts), VERTEX)
    for i, (p, n, t) in enumerate(verts):
        varr[i]["pos"] = p; varr[i]["norm"] = n; varr[i]["uv"] = t
    mats = np.array(mats_list, MATERIAL) if mats_list else np.zeros(0, MATERIAL)
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
