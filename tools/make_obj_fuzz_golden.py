#!/usr/bin/env python
"""tests/golden/obj_fuzz_streams.json: digests of what the REFERENCE's loader (vendored tinyobjloader v1.1.0 compiled from the
reference tree, oracle/_ref/tinyobj_dump, called as src/renderer.cpp:417 does) reads from the seeded random OBJ / MTL files of
tests/obj_fuzz.py.  Needs the reference tree (make -C oracle ref); the test suite then checks both of our readers against these
digests on any box.

    python tools/make_obj_fuzz_golden.py [n_seeds]
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import obj_fuzz  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    if not os.path.exists(obj_fuzz.TINYOBJ_DUMP):
        sys.exit("oracle/_ref/tinyobj_dump is missing: make -C oracle ref (needs /root/reference)")
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for seed in range(n):
            out[str(seed)] = obj_fuzz.digest(obj_fuzz.tinyobj_streams(obj_fuzz.write_case(d, seed)))
    path = os.path.join(ROOT, "tests", "golden", "obj_fuzz_streams.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/obj_fuzz.py write_case(seed)", "loader": "tinyobjloader v1.1.0 (reference tree), triangulate = true",
                   "cases": out}, f, indent=0)
    print(f"{n} cases, {sum(v == 'no model' for v in out.values())} without a model -> {path}")


if __name__ == "__main__":
    main()
