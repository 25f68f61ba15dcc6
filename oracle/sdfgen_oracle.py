"""ctypes front-end of oracle/sdfgen_oracle.c (CPU restatement of the reference's sdfgen).

TEST INFRASTRUCTURE ONLY - imported by tests/ and bench legs that check / time the CPU path,
never by the product package. Parity status: pinned against the reference's own
sdfgen::make_level_set3 (oracle/_ref/ref_probe.make_level_set3, tests/golden/sdfgen_small.npz).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsdfgen_oracle.so")
_SRC = os.path.join(_HERE, "sdfgen_oracle.c")
_lib = None


def build(force=False):
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c99", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.sdfgen_make_level_set3.restype = C.c_int
        _lib.sdfgen_make_level_set3.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                                C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.sdfgen_grid_layout.restype = None
        _lib.sdfgen_grid_layout.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]
    return _lib


def make_level_set3(tri, x, origin, dx, dy, dz, ni, nj, nk, exact_band=1, want_closest=False):
    """-> phi (ni*nj*nk,) float32, x fastest [, closest triangle per cell]."""
    tri = np.ascontiguousarray(tri, dtype=np.uint32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    origin = np.ascontiguousarray(origin, dtype=np.float32)
    phi = np.empty(ni * nj * nk, dtype=np.float32)
    closest = np.empty(ni * nj * nk, dtype=np.int32) if want_closest else None
    rc = lib().sdfgen_make_level_set3(len(tri), tri.ctypes.data, len(x), x.ctypes.data, origin.ctypes.data, np.float32(dx),
                                      np.float32(dy), np.float32(dz), ni, nj, nk, exact_band, phi.ctypes.data,
                                      closest.ctypes.data if want_closest else None)
    if rc != 0:
        raise MemoryError("sdfgen_make_level_set3")
    return (phi, closest) if want_closest else phi


def grid_layout(V, dx, res=20):
    """BodySDFObj::precompute_SDF layout (BodySDFObj.cpp:38-70) -> dict(spacing, min_box, max_box, dims)."""
    V = np.ascontiguousarray(V, dtype=np.float64)
    spacing = np.empty(3)
    mn, mx = np.empty(3, dtype=np.float32), np.empty(3, dtype=np.float32)
    dims = np.empty(3, dtype=np.int32)
    lib().sdfgen_grid_layout(len(V), V.ctypes.data, float(dx), int(res), spacing.ctypes.data, mn.ctypes.data, mx.ctypes.data,
                             dims.ctypes.data)
    return {"spacing": spacing, "min_box": mn, "max_box": mx, "dims": dims}
