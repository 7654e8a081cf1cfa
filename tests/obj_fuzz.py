"""Seeded random OBJ / MTL files for the differential test of the OBJ readers (tests/test_obj_reader_fuzz.py):
the reference parses its models with the vendored tinyobjloader v1.1.0 (src/renderer.cpp:417, triangulate = true);
oracle/_ref/tinyobj_dump is that loader compiled from the reference tree.  The generator stays inside what an
exporter writes (Blender / the Cornell box files: v, vn, vt, f, g, o, s, usemtl, mtllib, comments) but mixes the
syntax freely: all four corner forms, relative indices, polygons, every number format tinyobj's own float parser
accepts (and a few it does not), CRLF files, statements in unusual order."""
from __future__ import annotations

import os
import random


ODD_NUMBERS = [".5", "-.5", "1e", "1e+", "5.", "abc", "1,5", "0x10", "inf", "nan", "1.5f", "--1", "+", "-", "1e400", "1e-400", "1.0e+2x", "007",
               "3.14159265358979323846", "0.1234567890123456789", "1E2", "2e-3.5", "1..2", "9999999999999999999999", "-0", "-0.0", "1e0"]
NASTY = False      # set per case by write_case: also write what an exporter would not


def _num(rng: random.Random, lo=-2.0, hi=2.0) -> str:
    if NASTY and rng.random() < 0.08:
        return rng.choice(ODD_NUMBERS)
    x = rng.uniform(lo, hi)
    k = rng.randrange(12)
    if k == 0: return "%d" % round(x)
    if k == 1: return "%.6f" % x
    if k == 2: return "%.4f" % x
    if k == 3: return "%g" % x
    if k == 4: return "%e" % x
    if k == 5: return "%.9f" % x
    if k == 6: return "%+.3f" % x
    if k == 7: return "%.3E" % x
    if k == 8: return "%.12g" % x
    if k == 9: return ("%.5f" % x).rstrip("0")          # "1." / "-0." forms
    if k == 10: return "%.2fe%+d" % (x, rng.randrange(-3, 3))
    return repr(x)


def _sep(rng: random.Random) -> str:
    return rng.choice([" ", " ", " ", "  ", "\t", " \t"])


def make_mtl(rng: random.Random, names) -> str:
    out = ["# fuzz materials", ""]
    if NASTY and rng.random() < 0.3:
        out += ["Kd 0.5 0.25 0.125", "Ns 7"]                # statements in front of the first newmtl
    if NASTY and rng.random() < 0.15:
        names = []                                         # a file without any newmtl
        out += ["Ka 0.1 0.2 0.3", "illum 3"]
    for n in names:
        sep = rng.choice([" ", " ", "\t", "  "]) if NASTY else " "
        tail = rng.choice(["", "", " ", "\t"]) if NASTY else ""
        out.append(rng.choice(["", "  ", "\t"]) + "newmtl" + sep + n + tail)
        keys = ["Ka", "Kd", "Ks", "Ke", "Tf", "Kt", "Ns", "Ni", "d", "Tr", "illum", "map_Kd", "Pr", "Pm", "bogus"]
        rng.shuffle(keys)
        for key in keys[: rng.randrange(0, len(keys) + 1)]:
            if key in ("Ka", "Kd", "Ks", "Ke", "Tf", "Kt"):
                out.append(f"{key}{_sep(rng)}" + _sep(rng).join(_num(rng, 0, 1) for _ in range(rng.choice([3, 3, 3, 1, 2]))))
            elif key == "illum":
                out.append("illum " + (rng.choice(["2.7", "abc", "-3", " 4", "5 6"]) if NASTY and rng.random() < 0.3 else str(rng.randrange(0, 8))))
            elif key == "map_Kd":
                out.append("map_Kd texture.png")
            elif key == "bogus":
                out.append("bogus 1 2 3")
            else:
                out.append(f"{key}{_sep(rng)}{_num(rng, 0, 2)}")
        if rng.random() < 0.3:
            out.append("# trailing comment")
        out.append("")
    return "\n".join(out) + "\n"


def make_obj(rng: random.Random, mtl_name, names) -> str:
    out = ["# fuzz object"]
    if mtl_name is not None:
        if NASTY and rng.random() < 0.2:
            out.append("mtllib nowhere.mtl " + mtl_name)       # several file names: the first that loads is taken
        else:
            out.append("mtllib " + mtl_name)
        if NASTY and rng.random() < 0.15:
            out.append("mtllib " + mtl_name)                   # loaded twice: the materials are appended again
    nv = nn = nt = 0
    have_faces = False
    for block in range(rng.randrange(1, 6)):
        # a batch of attributes
        for _ in range(rng.randrange(3, 10)):
            extra = rng.random()
            line = "v" + _sep(rng) + _sep(rng).join(_num(rng) for _ in range(3))
            if extra < 0.1: line += " " + _num(rng, 0.5, 1.5)                      # w
            elif extra < 0.2: line += " " + " ".join(_num(rng, 0, 1) for _ in range(3))  # vertex colour
            out.append(rng.choice(["", "", " ", "\t"]) + line + rng.choice(["", "", " ", "  "]))
            nv += 1
        for _ in range(rng.randrange(0, 5)):
            out.append("vn" + _sep(rng) + _sep(rng).join(_num(rng, -1, 1) for _ in range(3))); nn += 1
        for _ in range(rng.randrange(0, 5)):
            out.append("vt" + _sep(rng) + _sep(rng).join(_num(rng, 0, 1) for _ in range(rng.choice([2, 2, 3, 1])))); nt += 1
        # statements between the attributes and the faces
        for _ in range(rng.randrange(0, 4)):
            k = rng.randrange(8)
            if k == 0: out.append("g " + rng.choice(["left", "right wall", "grp%d" % block]))
            elif k == 1: out.append("o obj%d" % block)
            elif k == 2: out.append("s " + rng.choice(["off", "1", "2"]))
            elif k == 3 and names:
                out.append("usemtl" + (rng.choice([" ", " ", "\t", "  "]) if NASTY else " ") + rng.choice(names)
                           + (rng.choice(["", "", " ", " # c"]) if NASTY else ""))
            elif k == 4: out.append("usemtl not_in_the_mtl")
            elif k == 5: out.append(rng.choice(["g", "o", "g ", "  g x", "usemtl"]) if NASTY else "g")
            elif k == 6: out.append("# comment" + rng.choice(["", " f 1 2 3"]))
            else: out.append("")
        for _ in range(rng.randrange(0, 8)):
            if rng.random() < 0.2 and names:
                out.append("usemtl " + rng.choice(names))
            if rng.random() < 0.07:
                out.append(rng.choice(["g mid", "o mid", "s 1"]))
            n = rng.choice([3, 3, 3, 4, 4, 5, 6])
            form = rng.randrange(4)
            if form in (1, 3) and nt == 0: form = 0
            if form in (2, 3) and nn == 0: form = 0 if form == 2 or nt == 0 else 1
            corners = []
            for _c in range(n):
                def ref(count):
                    i = rng.randrange(1, count + 1)
                    return str(i) if rng.random() < 0.7 else str(i - count - 1)     # relative form of the same element
                v = ref(nv)
                if form == 0: corners.append(v)
                elif form == 1: corners.append(v + "/" + ref(nt))
                elif form == 2: corners.append(v + "//" + ref(nn))
                else: corners.append(v + "/" + ref(nt) + "/" + ref(nn))
            if NASTY and rng.random() < 0.05:
                corners = corners[:2]                              # a face of two corners: no triangle, but a non-empty group
            if NASTY and rng.random() < 0.02:
                corners[rng.randrange(len(corners))] = rng.choice(["0", "x", "2/0", "1//0"])  # the load fails
            out.append("f" + _sep(rng) + _sep(rng).join(corners) + rng.choice(["", "", " "]))
            have_faces = True
    if rng.random() < 0.15 and names:
        out.append("usemtl " + rng.choice(names))       # file ends with a usemtl (tinyobj #104)
    if not have_faces:
        out.append("f 1 2 3")
    return "\n".join(out) + "\n"


def write_case(directory: str, seed: int):
    """Writes fuzz_<seed>.obj (+ .mtl) into `directory`; returns the OBJ path."""
    global NASTY
    rng = random.Random(0xB200 + seed)
    NASTY = seed % 3 == 2
    names = ["m%d" % i for i in range(rng.randrange(0, 5))]
    if NASTY and names and rng.random() < 0.3:
        names[-1] = "two words"
    if NASTY and len(names) > 1 and rng.random() < 0.2:
        names[1] = names[0]                                   # the same name twice: the first keeps it
    mode = rng.randrange(6)
    mtl_name = None if (mode == 0 or not names) else ("fuzz_%d.mtl" % seed)
    if mode == 1:
        mtl_name = "missing_%d.mtl" % seed                 # mtllib names a file that does not exist
    obj = make_obj(rng, mtl_name, names)
    eol = "\r\n" if rng.random() < 0.25 else ("\r" if NASTY and rng.random() < 0.1 else "\n")
    path = os.path.join(directory, "fuzz_%d.obj" % seed)
    with open(path, "w", newline="") as f:
        f.write(obj.replace("\n", eol))
    if mtl_name is not None and mode != 1:
        with open(os.path.join(directory, mtl_name), "w", newline="") as f:
            f.write(make_mtl(rng, names).replace("\n", eol))
    return path


# ----------------------------------------------------------------------------- streams and digests
# What is compared: the flattened per-index stream the reference's draw calls see -- position, normal, texcoord (float bit patterns)
# and the material id of the face (-1 = none) -- after Renderer::load_model's vertex dedupe (src/renderer.cpp:500-507: equal vertices
# share the first-seen one; -0.0 == 0.0 there), plus the material constants create_material forwards.
import hashlib
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINYOBJ_DUMP = os.path.join(ROOT, "oracle", "_ref", "tinyobj_dump")          # the reference's vendored loader, compiled where it lies
CPP_DUMP = os.path.join(ROOT, "voxel_cone_tracing_b200", "host", "obj_dump")  # vct::load_obj (CPU only)


def _first_seen(bits: np.ndarray) -> np.ndarray:
    """load_model's dedupe applied to a raw per-index stream: rows that compare equal as floats take the first one's bits."""
    out = bits.copy()
    seen = {}
    for i in range(len(bits)):
        key = tuple(0 if b == 0x80000000 else int(b) for b in bits[i])
        j = seen.setdefault(key, i)
        out[i] = bits[j]
    return out


def tinyobj_streams(obj_path: str):
    """(vertex stream u32[n, 9], materials u32[m, 19]) from the reference's loader, or None if LoadObj fails."""
    r = subprocess.run([TINYOBJ_DUMP, obj_path], capture_output=True, text=True)
    if r.returncode:
        return None
    V, M = [], []
    for line in r.stdout.splitlines():
        t = line.split()
        if t[0] == "V":
            V.append([int(x, 16) for x in t[1:9]] + [int(t[9]) & 0xFFFFFFFF])
        elif t[0] == "M":
            M.append([int(x, 16) for x in t[2:20]] + [int(t[20]) & 0xFFFFFFFF])
    V = np.array(V, np.uint32).reshape(-1, 9)
    if len(V):
        V[:, :8] = _first_seen(V[:, :8])
    return V, np.array(M, np.uint32).reshape(-1, 19)


def mesh_streams(mesh):
    """The same two arrays from one of our Mesh objects."""
    mat = np.full(len(mesh.indices), -1, np.int64)
    for (a, b, m) in mesh.ranges:
        mat[a:a + b] = m
    v = mesh.verts[mesh.indices]
    bits = np.concatenate([v["pos"], v["norm"], v["uv"]], axis=1).astype("<f4").view("<u4")
    V = np.concatenate([bits, (mat & 0xFFFFFFFF).astype(np.uint32)[:, None]], axis=1)
    M = []
    for m in mesh.materials:
        f = np.concatenate([m["ambient"][:3], m["diffuse"][:3], m["specular"][:3], m["transmittance"][:3], m["emission"],
                            [m["shininess"], m["ior"], m["dissolve"]]]).astype("<f4").view("<u4")
        M.append([int(x) for x in f] + [int(m["illum"]) & 0xFFFFFFFF])
    return V, np.array(M, np.uint32).reshape(-1, 19)


def digest(streams) -> str:
    """One string per case; "no model" where the load fails or yields no triangle (the reference prints an error or builds an empty model)."""
    if streams is None or len(streams[0]) == 0:
        return "no model"
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(streams[0], "<u4").tobytes())
    h.update(b"|")
    h.update(np.ascontiguousarray(streams[1], "<u4").tobytes())
    return f"{len(streams[0])}:{len(streams[1])}:{h.hexdigest()[:32]}"
