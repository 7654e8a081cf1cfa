#!/usr/bin/env python
"""glsl2cpp.py -- turns the REFERENCE'S OWN shader text into C++ that compiles against the reference's vendored GLM.

TEST INFRASTRUCTURE ONLY (oracle/).  Nothing is copied into the repository: the shaders are read from where they lie
(/root/reference/shader/*.vert|geom|frag|comp) at build time; the generated text lives in oracle/_ref/gen/ (git-ignored) only while the library is being compiled.

The translation is purely syntactic -- no expression, constant or statement of a shader body is touched:
  * `#version`, `layout(...)` qualifiers and the `uniform` / `in` / `out` storage qualifiers are dropped, so every global
    of the shader becomes a data member of the C++ struct the harness wraps around the generated text;
  * uniform blocks (`uniform camera { ... };`) are flattened into their members, interface blocks
    (`in GS_OUT { ... } gs_out;`) become plain structs, the geometry shader's unsized input array gets its size 3;
  * GLSL arrays (`vec4 values[8]`, `vec4[8] f()`, `ivec3[8](...)`) become glsl_array<T, N> (std::array): C++ cannot return
    or assign built-in arrays;
  * globals with a non-constant initialiser (`float voxel_size = 1.0f / cube_res;`) are declared without it and the
    initialisers are collected, in order, into `void _init_globals()`, which the harness calls before every `main()` --
    GLSL evaluates them at the start of each invocation;
  * unsuffixed floating literals get the `f` suffix (`0.04` is a float in GLSL, a double in C++);
  * every `#define` of the shader is `#undef`-ed at the end so the shaders can share one translation unit.
What GLSL means beyond C++ syntax (swizzles, implicit int->float conversions, built-ins, samplers, images, EmitVertex)
is supplied by oracle/glsl_ref/glsl_env.h.

    python glsl2cpp.py <shader dir> <out dir>
"""
from __future__ import annotations

import os
import re
import sys

SHADERS = ["voxelize.vert", "voxelize.geom", "voxelize.frag", "mipmap.comp", "voxel_cone_tracing.vert", "voxel_cone_tracing.frag"]
TYPES = r"(?:vec2|vec3|vec4|ivec3|uvec3|float|int|uint|bool|sampler3D|image3D|uimage3D|point_light)"


def _match_paren(s: str, i: int) -> int:
    """index of the ')' matching the '(' at s[i]"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


def translate(src: str, stage: str) -> str:
    s = src
    s = re.sub(r"^\s*#version.*$", "", s, flags=re.M)
    # GLSL floating literals are single precision
    s = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)(?![\w.])", r"\1f", s)
    # stand-alone layout declarations: `layout (triangles) in;`, `layout (local_size_x = 8, ...) in;`
    s = re.sub(r"^\s*layout\s*\([^)]*\)\s*(?:in|out)\s*;[^\n]*$", "", s, flags=re.M)
    # uniform blocks -> their members
    s = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+\w+\s*\{(.*?)\}\s*;", lambda m: m.group(1), s, flags=re.S)
    # interface blocks -> structs (+ instance); unsized geometry-shader input arrays hold one triangle

    def iface(m):
        arr = "[3]" if m.group(5) else ""
        return "struct %s {%s} %s%s;" % (m.group(2), m.group(3), m.group(4), arr)
    s = re.sub(r"\b(in|out)\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*(\[\s*\])?\s*;", iface, s, flags=re.S)
    # remaining layout qualifiers (also inside parameter lists), then storage qualifiers of globals
    s = re.sub(r"layout\s*\([^)]*\)", "", s)
    for _ in range(2):  # `uniform layout(..) T x;` leaves `uniform  T x;`; `layout(..) in T x;` leaves ` in T x;`
        s = re.sub(r"^(\s*)(?:uniform|in|out)\s+(?=\w)", r"\1", s, flags=re.M)
    # array constructors  T[N](a, b, ...)  ->  glsl_array<T, N>{{a, b, ...}}
    while True:
        m = re.search(r"\b(%s)\s*\[\s*(\d+)\s*\]\s*\(" % TYPES, s)
        if not m:
            break
        close = _match_paren(s, m.end() - 1)
        s = s[:m.start()] + "glsl_array<%s, %s>{{" % (m.group(1), m.group(2)) + s[m.end():close] + "}}" + s[close + 1:]
    # functions returning arrays  T[N] f(  ->  glsl_array<T, N> f(
    s = re.sub(r"\b(%s)\s*\[\s*(\w+)\s*\]\s+(\w+)\s*\(" % TYPES, r"glsl_array<\1, \2> \3(", s)
    # array declarations  T name[N]  ->  glsl_array<T, N> name   (unsized: size taken from the initialiser)

    def arr_decl(m):
        n = m.group(3).strip()
        if not n:
            init = re.match(r"\s*=\s*glsl_array<\w+,\s*(\d+)>", s[m.end():])
            if not init:
                raise ValueError("unsized array without constructor: " + m.group(0))
            n = init.group(1)
        return "glsl_array<%s, %s> %s" % (m.group(1), n, m.group(2))
    s = re.sub(r"\b(%s)\s+(\w+)\s*\[([^\]]*)\]" % TYPES, arr_decl, s)

    # globals with non-constant initialisers -> declaration + _init_globals()
    inits = []
    out, depth, pos = [], 0, 0
    for m in re.finditer(r"^(%s)\s+(\w+)\s*=\s*([^;]+);" % TYPES, s, flags=re.M):
        depth = s[:m.start()].count("{") - s[:m.start()].count("}")
        if depth != 0:
            continue
        out.append(s[pos:m.start()])
        out.append("%s %s;" % (m.group(1), m.group(2)))
        inits.append("  %s = %s;" % (m.group(2), m.group(3).strip()))
        pos = m.end()
    out.append(s[pos:])
    s = "".join(out)
    s += "\nvoid _init_globals()\n{\n" + "\n".join(inits) + "\n}\n"
    for name in re.findall(r"^\s*#define\s+(\w+)", s, flags=re.M):
        s += "#undef %s\n" % name
    return "// GENERATED by oracle/glsl_ref/glsl2cpp.py from the reference's shader/%s -- do not commit\n" % stage + s


def main(argv):
    src_dir, out_dir = argv[1], argv[2]
    os.makedirs(out_dir, exist_ok=True)
    for name in SHADERS:
        with open(os.path.join(src_dir, name)) as f:
            text = translate(f.read(), name)
        with open(os.path.join(out_dir, name.replace(".", "_") + ".inc"), "w") as f:
            f.write(text)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
