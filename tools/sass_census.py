#!/usr/bin/env python
"""Static census of the built library (no GPU needed): per kernel the ptxas resource usage (cuobjdump -res-usage)
and the count of the SASS mnemonics that say which hardware path a kernel uses -- TEX (texture units), SUST / SULD
(surface stores / loads of the mip array), LDGSTS (cp.async), ATOM / ATOMG / RED (atomics), LDG.E.128 / STG.E.128
(16-byte global accesses), MUFU, BAR, and the instruction total.

    python tools/sass_census.py [voxel_cone_tracing_b200/libvct_cuda.so] > profiles/r02_sass_census.md
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "voxel_cone_tracing_b200", "libvct_cuda.so")
CUOBJDUMP = "/usr/local/cuda/bin/cuobjdump"
FILT = "/usr/local/cuda/bin/cu++filt"

CLASSES = [
    ("TEX", re.compile(r"^TEX|^TLD")),
    ("SUST", re.compile(r"^SUST")),
    ("SULD", re.compile(r"^SULD")),
    ("LDGSTS", re.compile(r"^LDGSTS")),
    ("ATOM/RED", re.compile(r"^ATOM|^RED|^ATOMG|^ATOMS")),
    ("LDG.128", re.compile(r"^LDG\..*128")),
    ("STG.128", re.compile(r"^STG\..*128")),
    ("MUFU", re.compile(r"^MUFU")),
    ("BAR", re.compile(r"^BAR")),
    ("SHFL/VOTE/MATCH", re.compile(r"^SHFL|^VOTE|^MATCH|^REDUX")),
]


def demangle(names):
    out = subprocess.run([FILT] + names, capture_output=True, text=True).stdout.split("\n")
    res = []
    for o in out[: len(names)]:
        o = o.replace("vct::", "").replace("void ", "")
        o = o[: o.rfind(">(") + 1] if ">(" in o else o.split("(")[0]
        res.append(o.replace("(bool)", "").replace("(int)", ""))
    return res


def main():
    res = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line))
            cur = None
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for name, rx in CLASSES:
                if rx.search(op):
                    counts[cur][name] += 1
    names = list(counts)
    pretty = dict(zip(names, demangle(names)))
    print("# Static census of `libvct_cuda.so` (sm_100a; `tools/sass_census.py`, cuobjdump -res-usage / -sass, no GPU involved)\n")
    print("Counts are static SASS instructions per kernel, not executed ones.  TEX = texture-unit fetches (cone tracer, point-sampled")
    print("downloads), SUST = surface stores into the mipmapped array, LDGSTS = `cp.async` (the mip kernel's slab ring), ATOM/RED = the")
    print("voxelizer's list push / counters, the G-buffer's 64-bit `atomicMin`, the work queues.  No tensor-core or TMA instruction is")
    print("expected on this path (nothing is a contraction; the mip ring moved from TMA to per-warp cp.async in round 2).")
    print("Template arguments: `cone_kernel_fast|grid<TEX sampler, SPLIT (one- / two-level fetches split), MIN_CTAS per SM, GROUP (all diffuse")
    print("cones of a tile in one warp)>` -- `<1, 1, 10, 1>` is the kernel of the benchmarked config 2 and of the 4K / 8K frames, `<1, 1, 10, 0>` the one of frames with fewer than 32768 tiles per rank (config 1, N >= 4);")
    print("`cone_kernel<COUNT, TEX, F16>` = the literal march (sample counters, RGBA16F variant); `cam_setup_kernel<SPLIT>` / `cam_resolve_kernel<LEAN>`:")
    print("frame shared between ranks / register-capped resolve.\n")
    hdr = ["kernel", "regs", "stack B", "smem B", "SASS"] + [c for c, _ in CLASSES]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    for n in sorted(names, key=lambda k: pretty[k]):
        u = usage.get(n, {})
        c = counts[n]
        row = [f"`{pretty[n]}`", u.get("REG", "?"), u.get("STACK", "?"), u.get("SHARED", "?"), str(c["total"])] + [str(c[k]) if c[k] else "" for k, _ in CLASSES]
        print("| " + " | ".join(row) + " |")
    all_ops = collections.Counter()
    for line in sass.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            all_ops[m.group(1)] += 1
    probe = ["UTMALDG", "UTMASTG", "UTCMMA", "UTCHMMA", "HMMA", "IMMA", "DMMA", "SYNCS", "LDGSTS", "TEX", "SUST", "ATOMG", "RED", "CCTL"]
    print("\nWhole-library mnemonic probe: " + ", ".join(f"{p} {all_ops.get(p, 0)}" for p in probe))


if __name__ == "__main__":
    main()
