"""CPU check of the arithmetic the mip kernels use (csrc/mip_arith.cuh, compiled by g++ into tests/native/libvct_hosttest.so):
the integer dot-product formulation of one anisotropic mip step (shader/mipmap.comp:22-100) + the fp32 replay of exact ties
must reproduce the oracle's fp32 recipe bit for bit -- on random data (ties are rare), on opaque / binary-alpha data (a quarter
of the channels tie) and on hand-picked tie patterns."""
import ctypes
import os

import numpy as np
import pytest

from oracle import orc

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "libvct_hosttest.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.fail(f"{LIB} is missing: run `make hosttest`")
    L = ctypes.CDLL(LIB)
    L.vct_hosttest_mip_step.restype = ctypes.c_uint64
    L.vct_hosttest_mip_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return L


def _grid(kind: str, R: int, rng) -> np.ndarray:
    base = rng.integers(0, 2 ** 32, (R, R, R), dtype=np.uint64).astype(np.uint32)
    if kind == "opaque":
        base |= np.uint32(0xFF000000)
        base[rng.random((R, R, R)) < 0.3] = 0xFFFFFFFF
    elif kind == "sparse":
        base[rng.random((R, R, R)) > 0.05] = 0
    elif kind == "ties":
        base = rng.choice(np.array([0, 0xFEFEFEFF, 0xFFFFFFFF, 0x80808080, 0x00FF00FE, 0xFE0000FE, 0x020202FE], np.uint32), (R, R, R))
    elif kind == "count_nibble":   # what the voxelizer stores: 7-bit colour + count bits in the LSBs, alpha 254/255
        base = (base & np.uint32(0x00FEFEFE)) | np.uint32(0xFE000000) | (rng.integers(0, 2, (R, R, R)).astype(np.uint32) * np.uint32(0x01010101))
        base[rng.random((R, R, R)) > 0.4] = 0
    return np.ascontiguousarray(base)


@pytest.mark.parametrize("kind", ["random", "opaque", "sparse", "ties", "count_nibble"])
def test_integer_mip_step_matches_oracle(lib, kind):
    R = 32
    rng = np.random.default_rng(3)
    base = _grid(kind, R, rng)
    levels = 6
    pyr = orc.mipmap(base, levels)
    dst = np.empty((6, R // 2, R // 2, R // 2), np.uint32)
    ties = lib.vct_hosttest_mip_step(base.ctypes.data, R, 1, dst.ctypes.data)
    for d in range(6):
        assert np.array_equal(dst[d], pyr.levels[d][1]), f"{kind}: level 1 direction {d}"
    if kind in ("opaque", "ties", "count_nibble"):
        assert ties > 0.01 * dst.size * 4, "this data is meant to exercise the tie replay"
    for l in range(1, levels - 1):
        N = R >> l
        src = np.ascontiguousarray(np.stack([pyr.levels[d][l] for d in range(6)]))
        dst = np.empty((6, N // 2, N // 2, N // 2), np.uint32)
        lib.vct_hosttest_mip_step(src.ctypes.data, N, 0, dst.ctypes.data)
        for d in range(6):
            assert np.array_equal(dst[d], pyr.levels[d][l + 1]), f"{kind}: level {l + 1} direction {d}"
