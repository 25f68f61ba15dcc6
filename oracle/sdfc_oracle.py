"""ctypes front-end of the CPU oracle (oracle/sdfc_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package.
Parity status: pinned against the reference itself (see sdfc_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsdfc_oracle.so")
_SRC = os.path.join(_HERE, "sdfc_oracle.c")


def build(force=False):
    """gcc the C restatement into oracle/libsdfc_oracle.so (no FMA contraction)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c99", "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


class _Grid(C.Structure):
    _fields_ = [("ni", C.c_int), ("nj", C.c_int), ("nk", C.c_int),
                ("min_box", C.c_double * 3), ("max_box", C.c_double * 3),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("col_th", C.c_double), ("grid", C.POINTER(C.c_double))]


class _Params(C.Structure):
    _fields_ = [("kn", C.c_double), ("kt", C.c_double), ("mu", C.c_double), ("damping", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.sdfo_query_sdf.restype = C.c_double
        _lib.sdfo_distance.restype = C.c_double
        _lib.sdfo_detect.restype = C.c_int
        _lib.sdfo_contact_term.restype = C.c_int
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _colmajor(E):
    """4x4 (or any 2-D) numpy matrix -> flat column-major fp64 (Eigen layout)."""
    return np.ascontiguousarray(np.asarray(E, dtype=np.float64).T).reshape(-1)


class Grid:
    """Voxel SDF of one body (fields of BodySDFObj, Body/BodySDFObj.h:13-16)."""

    def __init__(self, ni, nj, nk, min_box, max_box, dx, dy, dz, col_th, grid):
        self.grid = np.ascontiguousarray(grid, dtype=np.float64).reshape(-1)
        assert self.grid.size == ni * nj * nk
        self.ni, self.nj, self.nk = int(ni), int(nj), int(nk)
        self.min_box = np.asarray(min_box, dtype=np.float64).copy()
        self.max_box = np.asarray(max_box, dtype=np.float64).copy()
        self.dx, self.dy, self.dz, self.col_th = float(dx), float(dy), float(dz), float(col_th)
        self._c = _Grid(self.ni, self.nj, self.nk, (C.c_double * 3)(*self.min_box), (C.c_double * 3)(*self.max_box),
                        self.dx, self.dy, self.dz, self.col_th, _dp(self.grid))

    @classmethod
    def from_dict(cls, d):
        return cls(d["ni"], d["nj"], d["nk"], d["min_box"], d["max_box"], d["dx"], d["dy"], d["dz"], d["col_th"], d["grid"])

    @classmethod
    def from_spec(cls, s):
        return cls(s.ni, s.nj, s.nk, s.min_box, s.max_box, s.dx, s.dy, s.dz, s.col_th, s.grid)

    def set_col_th(self, col_th):
        self.col_th = float(col_th)
        self._c.col_th = self.col_th

    @property
    def ref(self):
        return C.byref(self._c)


def _pts(X):
    X = np.ascontiguousarray(X, dtype=np.float64)
    assert X.ndim == 2 and X.shape[1] == 3
    return X


def query_sdf(g, X, outside_accurate=False):
    X = _pts(X)
    out = np.empty(len(X))
    lib().sdfo_query_sdf_batch(g.ref, len(X), _dp(X), int(outside_accurate), _dp(out))
    return out


def query_dsdf(g, X):
    X = _pts(X)
    out = np.empty((len(X), 3))
    lib().sdfo_query_dsdf_batch(g.ref, len(X), _dp(X), _dp(out))
    return out


def query_ddsdf(g, X):
    X = _pts(X)
    out = np.empty((len(X), 3, 3))
    lib().sdfo_query_ddsdf_batch(g.ref, len(X), _dp(X), _dp(out))
    return out


def grid_location(g, X):
    X = _pts(X)
    out = np.empty((len(X), 3), dtype=np.int32)
    lib().sdfo_grid_location_batch(g.ref, len(X), _dp(X), _ip(out))
    return out


def mesh_distance(g, X):
    X = _pts(X)
    out = np.empty(len(X))
    lib().sdfo_mesh_distance_batch(g.ref, len(X), _dp(X), _dp(out))
    return out


def detect(g, E1, E2, pts):
    """-> dict(d[n], cell[n,3], xw[n,3], active_ids[k]) in point order."""
    pts = _pts(pts)
    n = len(pts)
    d = np.empty(n)
    cell = np.empty((n, 3), dtype=np.int32)
    xw = np.empty((n, 3))
    ids = np.empty(n, dtype=np.int32)
    e1, e2 = _colmajor(E1), _colmajor(E2)
    k = lib().sdfo_detect(g.ref, _dp(e1), _dp(e2), n, _dp(pts), _dp(d), _ip(cell), _dp(xw), _ip(ids))
    return {"d": d, "cell": cell, "xw": xw, "active_ids": ids[:k].copy()}


def pair_force(g, E1, phi1, E2, phi2, pts, kn, kt, mu, damping, deriv=True):
    """One (contact body, SDF body) pair. kn/kt already include scale*contact_scale.
    -> fm[12], Km[12,12], Dm[12,12] (None if not deriv), (n_active, n_static)."""
    pts = _pts(pts)
    p = _Params(kn, kt, mu, damping)
    e1, e2 = _colmajor(E1), _colmajor(E2)
    f1 = np.ascontiguousarray(phi1, dtype=np.float64)
    f2 = np.ascontiguousarray(phi2, dtype=np.float64)
    fm = np.zeros(12)
    Km = np.zeros(144)
    Dm = np.zeros(144)
    counts = np.zeros(2, dtype=np.int32)
    lib().sdfo_pair_force(g.ref, C.byref(p), _dp(e1), _dp(f1), _dp(e2), _dp(f2), len(pts), _dp(pts), int(deriv),
                          _dp(fm), _dp(Km), _dp(Dm), _ip(counts))
    if not deriv:
        return fm, None, None, tuple(int(c) for c in counts)
    return fm, Km.reshape(12, 12).T.copy(), Dm.reshape(12, 12).T.copy(), tuple(int(c) for c in counts)


def contact_term(g, E1, phi1, E2, phi2, xi1, kn, kt, mu, damping, deriv=True):
    p = _Params(kn, kt, mu, damping)
    e1, e2 = _colmajor(E1), _colmajor(E2)
    f1 = np.ascontiguousarray(phi1, dtype=np.float64)
    f2 = np.ascontiguousarray(phi2, dtype=np.float64)
    x = np.ascontiguousarray(xi1, dtype=np.float64)
    fm = np.zeros(12)
    Km = np.zeros(144)
    Dm = np.zeros(144)
    st = lib().sdfo_contact_term(g.ref, C.byref(p), _dp(e1), _dp(f1), _dp(e2), _dp(f2), _dp(x), int(deriv),
                                 _dp(fm), _dp(Km), _dp(Dm))
    return fm, Km.reshape(12, 12).T.copy(), Dm.reshape(12, 12).T.copy(), int(st)


def scatter(ndof_m, idx1, idx2, fm, Km, Dm, fm_sys, Km_sys, Dm_sys):
    """Accumulate pair blocks into system arrays (numpy, same index arithmetic as
    ForceGeneralSDFContact.cpp:60-61,88-99)."""
    idx = [idx1, idx2]
    for b in range(2):
        fm_sys[idx[b]:idx[b] + 6] += fm[6 * b:6 * b + 6]
    if Km is not None:
        for br in range(2):
            for bc in range(2):
                Km_sys[idx[br]:idx[br] + 6, idx[bc]:idx[bc] + 6] += Km[6 * br:6 * br + 6, 6 * bc:6 * bc + 6]
                Dm_sys[idx[br]:idx[br] + 6, idx[bc]:idx[bc] + 6] += Dm[6 * br:6 * br + 6, 6 * bc:6 * bc + 6]
