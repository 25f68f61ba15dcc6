"""Host-side mirror of the reference's SDF-contact plug-in, over the C-ABI.

Names and semantics follow the reference (paths relative to
/root/reference/simulation/redmax):

  SdfGrid                  <- BodySDFObj's grid fields (Body/BodySDFObj.h:13-16) and
                              SDFMesh.distance / min_distance (Mesh.cpp:143-198)
  ContactPoints            <- Body::_contact_points (Body/Body.h:36)
  ForceGeneralSDFContact   <- Force/ForceGeneralSDFContact.{h,cpp}: set_stiffness /
                              set_friction / set_damping / set_scale, computeForce,
                              computeForceWithDerivative with ACCUMULATE (+=) semantics
                              into caller-owned fm / Km / Dm
  ContactBatch             <- Robot::computeForce[WithDerivative]'s loop over _forces
                              (Robot.cpp:513-551) for E environments at once

All arithmetic runs in the CUDA library; this module only marshals numpy arrays.
Poses are numpy (4,4) matrices (converted to Eigen's column-major layout here).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SdfContactError, check  # noqa: F401


def _dp(a):
    return a.ctypes.data_as(_lib.c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor44(E):
    E = np.asarray(E, dtype=np.float64)
    assert E.shape[-2:] == (4, 4)
    return np.ascontiguousarray(np.swapaxes(E, -1, -2))


class PinnedArray:
    """numpy array over page-locked host memory (sdfc_pinned_alloc): what the host-buffer
    entries DMA from / into without a staging copy. Keep the object alive while `.a` is used."""

    def __init__(self, shape, dtype=np.float64):
        self._lib = _lib.load()
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        self._p = C.c_void_p()
        check(self._lib.sdfc_pinned_alloc(n, C.byref(self._p)))
        buf = (C.c_char * max(n, 1)).from_address(self._p.value)
        self.a = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
        self.ptr = int(self._p.value)   # address of a[0]: eval_into takes it without building a ctypes pointer

    def close(self):
        if self._p:
            self.a = None
            self._lib.sdfc_pinned_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    def __init__(self, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        check(self._lib.sdfc_context_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def set_stream(self, cuda_stream_ptr):
        check(self._lib.sdfc_context_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        check(self._lib.sdfc_context_synchronize(self._h))

    @property
    def launch_count(self):
        return int(self._lib.sdfc_context_launch_count(self._h))

    KERNEL_SLOTS = ("detect", "force", "reduce", "scatter")

    def enable_timing(self, enable=True):
        check(self._lib.sdfc_context_enable_timing(self._h, int(bool(enable))))

    def read_timing(self, reset=True):
        """-> {slot: (total_ms, launches)} accumulated since the last reset (synchronises)."""
        ms = (C.c_double * 4)()
        cnt = (C.c_int64 * 4)()
        check(self._lib.sdfc_context_read_timing(self._h, ms, cnt, int(bool(reset))))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNEL_SLOTS)}

    def measure_peaks(self):
        """{"fp64_gflops", "l2_read_gbs"}: DFMA and L2 -> SM ceilings measured on this device (roofline denominators)."""
        out = (C.c_double * 4)()
        check(self._lib.sdfc_measure_peaks(self._h, out))
        return {"fp64_gflops": float(out[0]), "l2_read_gbs": float(out[1])}

    def selftest_exact_div(self, n, seed, b):
        """Mismatches between the library's reciprocal-based division and IEEE `/` (must be 0)."""
        bad = C.c_int64(-1)
        check(self._lib.sdfc_selftest_exact_div(self._h, int(n), int(seed), float(b), C.byref(bad)))
        return int(bad.value)

    def close(self):
        if self._h:
            self._lib.sdfc_context_destroy(self._h)
            self._h = C.c_void_p()


class SdfGrid:
    def __init__(self, ctx, ni, nj, nk, min_box, max_box, dx, dy, dz, col_th, grid):
        self.ctx = ctx
        self._lib = ctx._lib
        g = _f64(np.asarray(grid).reshape(-1))
        if g.size != int(ni) * int(nj) * int(nk):
            raise SdfContactError("[Error] SDF grid size does not match ni*nj*nk")
        mn, mx = _f64(min_box), _f64(max_box)
        self.ni, self.nj, self.nk = int(ni), int(nj), int(nk)
        self.min_box, self.max_box = mn.copy(), mx.copy()
        self.dx, self.dy, self.dz, self.col_th = float(dx), float(dy), float(dz), float(col_th)
        self._h = C.c_void_p()
        check(self._lib.sdfc_sdf_create(ctx._h, self.ni, self.nj, self.nk, _dp(mn), _dp(mx), self.dx, self.dy, self.dz,
                                        self.col_th, _dp(g), C.byref(self._h)))

    @classmethod
    def from_spec(cls, ctx, s):
        return cls(ctx, s.ni, s.nj, s.nk, s.min_box, s.max_box, s.dx, s.dy, s.dz, s.col_th, s.grid)

    @classmethod
    def from_dict(cls, ctx, d):
        return cls(ctx, d["ni"], d["nj"], d["nk"], d["min_box"], d["max_box"], d["dx"], d["dy"], d["dz"], d["col_th"],
                   d["grid"])

    @property
    def is_fp32(self):
        return bool(self._lib.sdfc_sdf_is_fp32(self._h))

    @property
    def layout(self):
        """Device storage: "fp64", "records" (fp32 cell records), "bricks" (fp32 4x4x4 bricks, TMA-staged by
        detection) or "bricks+records" (large grids: bricks for detection, records for per-point gathers)."""
        return ("fp64", "records", "bricks", "bricks+records")[int(self._lib.sdfc_sdf_layout_kind(self._h))]

    def set_col_th(self, col_th):
        self.col_th = float(col_th)
        check(self._lib.sdfc_sdf_set_col_th(self._h, self.col_th))

    def distance(self, X):
        """SDFMesh.distance: X is (n,3) in the grid's frame -> (n,) raw SDF (FLT_MAX outside)."""
        X = _f64(X)
        if X.ndim != 2 or X.shape[1] != 3:
            raise SdfContactError("[Error] distance expects an (n,3) array")
        out = np.empty(len(X), dtype=np.float64)
        check(self._lib.sdfc_sdf_distance(self._h, len(X), _dp(X), _dp(out)))
        return out

    def min_distance(self, X):
        X = _f64(X)
        if X.ndim != 2 or X.shape[1] != 3:
            raise SdfContactError("[Error] min_distance expects an (n,3) array")
        out = C.c_double()
        check(self._lib.sdfc_sdf_min_distance(self._h, len(X), _dp(X), C.byref(out)))
        return float(out.value)

    def stencil(self, X, fast=True):
        """BodySDFObj::query_SDF(x, true), query_dSDF, query_ddSDF (BodySDFObj.cpp:123-205) at X (n,3) in
        the grid's frame -> (sdf (n,), grad (n,3), hess (n,3,3)). fast selects the planner kernel's
        shared-stencil path, else the general kernel's sample-by-sample path; same bits either way."""
        X = _f64(X)
        if X.ndim != 2 or X.shape[1] != 3:
            raise SdfContactError("[Error] stencil expects an (n,3) array")
        n = len(X)
        acc, g, H = np.empty(n), np.empty((n, 3)), np.empty((n, 3, 3))
        check(self._lib.sdfc_sdf_stencil(self._h, n, _dp(X), int(bool(fast)), _dp(acc), _dp(g), _dp(H)))
        return acc, g, H

    def distance_device(self, n, dx_ptr, dy_ptr, dz_ptr, dout_ptr):
        check(self._lib.sdfc_sdf_distance_device(self._h, int(n), dx_ptr, dy_ptr, dz_ptr, dout_ptr))

    def close(self):
        if self._h:
            self._lib.sdfc_sdf_destroy(self._h)
            self._h = C.c_void_p()


def make_level_set3(ctx, tri, x, origin, dx, dy, dz, ni, nj, nk, exact_band=1):
    """sdfgen::make_level_set3 on the GPU -> phi (ni*nj*nk,) float32, x fastest."""
    tri = np.ascontiguousarray(tri, dtype=np.uint32).reshape(-1, 3)
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3)
    origin = np.ascontiguousarray(origin, dtype=np.float32)
    phi = np.empty(int(ni) * int(nj) * int(nk), dtype=np.float32)
    check(ctx._lib.sdfc_make_level_set3(ctx._h, len(tri), tri.ctypes.data, len(x), x.ctypes.data, origin.ctypes.data,
                                        np.float32(dx), np.float32(dy), np.float32(dz), int(ni), int(nj), int(nk),
                                        int(exact_band), phi.ctypes.data))
    return phi


def sdf_layout(V, dx, res=20):
    """BodySDFObj::precompute_SDF's grid layout -> dict(spacing (3,) f64, min_box, max_box (3,) f32, dims (3,) i32)."""
    lib = _lib.load()
    V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1, 3)
    spacing = np.empty(3)
    mn, mx = np.empty(3, dtype=np.float32), np.empty(3, dtype=np.float32)
    dims = np.empty(3, dtype=np.int32)
    check(lib.sdfc_sdf_layout(len(V), _dp(V), float(dx), int(res), _dp(spacing), mn.ctypes.data, mx.ctypes.data,
                              dims.ctypes.data_as(_lib.c_int32_p)))
    return {"spacing": spacing, "min_box": mn, "max_box": mx, "dims": dims}


def precompute_sdf(ctx, V, F, dx, col_th=0.0, res=20):
    """BodySDFObj::precompute_SDF (BodySDFObj.cpp:38-84) on the GPU: body-frame vertices V (n,3)
    fp64 and faces F (m,3) -> grid dict in the layout of SdfGrid.from_dict (fp64 values that are
    exactly the fp32 sdfgen output widened, as the reference stores them)."""
    lay = sdf_layout(V, dx, res)
    ni, nj, nk = (int(v) for v in lay["dims"])
    sp = lay["spacing"]
    phi = make_level_set3(ctx, F, np.asarray(V, dtype=np.float64).astype(np.float32), lay["min_box"], np.float32(sp[0]),
                          np.float32(sp[1]), np.float32(sp[2]), ni, nj, nk)
    return {"ni": ni, "nj": nj, "nk": nk, "min_box": lay["min_box"].astype(np.float64),
            "max_box": lay["max_box"].astype(np.float64), "dx": float(sp[0]), "dy": float(sp[1]), "dz": float(sp[2]),
            "col_th": float(col_th), "grid": phi.astype(np.float64)}


def save_sdf_text(path, g):
    """BodySDFObj::save_SDF: the `.sdf` cache next to the OBJ (6 significant digits)."""
    lib = _lib.load()
    grid = _f64(g["grid"]).reshape(-1)
    check(lib.sdfc_sdf_save_text(str(path).encode(), int(g["ni"]), int(g["nj"]), int(g["nk"]), _dp(_f64(g["min_box"])),
                                 _dp(_f64(g["max_box"])), float(g["dx"]), float(g["dy"]), float(g["dz"]), _dp(grid)))


def load_sdf_text(path, col_th=0.0):
    """BodySDFObj::load_SDF -> grid dict."""
    lib = _lib.load()
    dims = np.empty(3, dtype=np.int32)
    check(lib.sdfc_sdf_load_text_header(str(path).encode(), dims.ctypes.data_as(_lib.c_int32_p)))
    n = int(dims[0]) * int(dims[1]) * int(dims[2])
    mn, mx, sp, grid = np.empty(3), np.empty(3), np.empty(3), np.empty(n)
    check(lib.sdfc_sdf_load_text(str(path).encode(), _dp(mn), _dp(mx), _dp(sp), _dp(grid), n))
    return {"ni": int(dims[0]), "nj": int(dims[1]), "nk": int(dims[2]), "min_box": mn, "max_box": mx, "dx": float(sp[0]),
            "dy": float(sp[1]), "dz": float(sp[2]), "col_th": float(col_th), "grid": grid}


class ContactPoints:
    SORT_MODES = {False: 0, True: 1, None: 0, "none": 0, "morton": 1, "kd": 2, 0: 0, 1: 1, 2: 2}

    def __init__(self, ctx, xyz, morton_sort=True):
        """morton_sort: False / True (30-bit Morton order) / "kd" (k-d order with box-shaped detection tiles)."""
        self.ctx = ctx
        self._lib = ctx._lib
        X = _f64(xyz)
        if X.ndim != 2 or X.shape[1] != 3:
            raise SdfContactError("[Error] contact points must be an (n,3) array")
        self.n = len(X)
        self._h = C.c_void_p()
        check(self._lib.sdfc_points_create(ctx._h, self.n, _dp(X), self.SORT_MODES[morton_sort], C.byref(self._h)))

    def close(self):
        if self._h:
            self._lib.sdfc_points_destroy(self._h)
            self._h = C.c_void_p()


class ForceGeneralSDFContact:
    """One (contact body, SDF body) contact force; see module docstring."""

    def __init__(self, ctx, contact_points, sdf, kn=1.0, kt=0.0, mu=0.0, damping=0.0, index_contact=0, index_sdf=6):
        if not isinstance(sdf, SdfGrid):
            # ForceGeneralSDFContact.cpp:17-20
            raise SdfContactError("[Error] The second body in Contact should be SDF body.")
        self.ctx, self._lib = ctx, ctx._lib
        self.contact_points, self.sdf = contact_points, sdf
        self._kn, self._kt, self._mu, self._damping, self._scale = float(kn), float(kt), float(mu), float(damping), 1.0
        self.index_contact, self.index_sdf = int(index_contact), int(index_sdf)
        self._h = C.c_void_p()
        check(self._lib.sdfc_pair_create(ctx._h, contact_points._h, sdf._h, self._kn, self._kt, self._mu, self._damping,
                                         C.byref(self._h)))
        self.touched = False  # contact history flag (CollisionDetection.cpp:143-146)

    # ---- setters (ForceGeneralSDFContact.h:28-31)
    def set_stiffness(self, kn, kt):
        self._kn, self._kt = float(kn), float(kt)
        check(self._lib.sdfc_pair_set_stiffness(self._h, self._kn, self._kt))

    def set_friction(self, mu):
        self._mu = float(mu)
        check(self._lib.sdfc_pair_set_friction(self._h, self._mu))

    def set_damping(self, damping):
        self._damping = float(damping)
        check(self._lib.sdfc_pair_set_damping(self._h, self._damping))

    def set_scale(self, scale):
        self._scale = float(scale)
        check(self._lib.sdfc_pair_set_scale(self._h, self._scale))

    def set_contact_scale(self, contact_scale):
        check(self._lib.sdfc_pair_set_contact_scale(self._h, float(contact_scale)))

    # ---- detection only
    def query(self, E1, E2, want_d=True, want_cell=True, want_active=True):
        n = self.contact_points.n
        e1, e2 = _colmajor44(E1), _colmajor44(E2)
        d = np.empty(n, dtype=np.float64) if want_d else None
        cell = np.empty((n, 3), dtype=np.int32) if want_cell else None
        act = np.empty(n, dtype=np.uint8) if want_active else None
        cnt = C.c_int32()
        check(self._lib.sdfc_pair_query(
            self._h, _dp(e1), _dp(e2), _dp(d) if want_d else None,
            cell.ctypes.data_as(_lib.c_int32_p) if want_cell else None,
            act.ctypes.data_as(_lib.c_uint8_p) if want_active else None, C.byref(cnt)))
        return {"d": d, "cell": cell, "active": act, "n_active": int(cnt.value)}

    # ---- pair blocks
    def force_sub(self, E1, phi1, E2, phi2):
        e1, e2, f1, f2 = _colmajor44(E1), _colmajor44(E2), _f64(phi1), _f64(phi2)
        fm = np.empty(12)
        cnt = C.c_int32()
        check(self._lib.sdfc_pair_force(self._h, _dp(e1), _dp(f1), _dp(e2), _dp(f2), _dp(fm), C.byref(cnt)))
        self.touched = self.touched or cnt.value > 0
        return fm, int(cnt.value)

    def force_deriv_sub(self, E1, phi1, E2, phi2):
        e1, e2, f1, f2 = _colmajor44(E1), _colmajor44(E2), _f64(phi1), _f64(phi2)
        fm, Km, Dm = np.empty(12), np.empty(144), np.empty(144)
        cnt = C.c_int32()
        check(self._lib.sdfc_pair_force_deriv(self._h, _dp(e1), _dp(f1), _dp(e2), _dp(f2), _dp(fm), _dp(Km), _dp(Dm),
                                              C.byref(cnt)))
        self.touched = self.touched or cnt.value > 0
        return fm, Km.reshape(12, 12).T.copy(), Dm.reshape(12, 12).T.copy(), int(cnt.value)

    # ---- the reference's public wrappers (ForceGeneralSDFContact.cpp:46-105): accumulate
    def computeForce(self, fm, E1, phi1, E2, phi2):
        sub, _ = self.force_sub(E1, phi1, E2, phi2)
        fm[self.index_contact:self.index_contact + 6] += sub[:6]
        fm[self.index_sdf:self.index_sdf + 6] += sub[6:]

    def computeForceWithDerivative(self, fm, Km, Dm, E1, phi1, E2, phi2):
        sub, K, D, _ = self.force_deriv_sub(E1, phi1, E2, phi2)
        ia, ib = self.index_contact, self.index_sdf
        fm[ia:ia + 6] += sub[:6]
        fm[ib:ib + 6] += sub[6:]
        for (r0, rs) in ((ia, 0), (ib, 6)):
            for (c0, cs) in ((ia, 0), (ib, 6)):
                Km[r0:r0 + 6, c0:c0 + 6] += K[rs:rs + 6, cs:cs + 6]
                Dm[r0:r0 + 6, c0:c0 + 6] += D[rs:rs + 6, cs:cs + 6]

    def close(self):
        if self._h:
            self._lib.sdfc_pair_destroy(self._h)
            self._h = C.c_void_p()


class ForceGroundContact:
    """ForceGroundContact (Force/ForceGroundContact.h): contact points of one body against
    the plane through E_g's origin with normal E_g[:3, 2]."""

    def __init__(self, ctx, contact_points, E_g, kn=1.0, kt=0.0, mu=0.0, damping=0.0, index_contact=0):
        self.ctx, self._lib = ctx, ctx._lib
        self.contact_points = contact_points
        self.index_contact = int(index_contact)
        self._h = C.c_void_p()
        check(self._lib.sdfc_ground_create(ctx._h, contact_points._h, _dp(_colmajor44(E_g)), float(kn), float(kt), float(mu),
                                           float(damping), C.byref(self._h)))
        self.touched = False  # body->add_contact_body("ground"), CollisionDetection.cpp:67

    def set_transform(self, E_g):
        check(self._lib.sdfc_ground_set_transform(self._h, _dp(_colmajor44(E_g))))

    def set_stiffness(self, kn, kt):
        check(self._lib.sdfc_ground_set_stiffness(self._h, float(kn), float(kt)))

    def set_friction(self, mu):
        check(self._lib.sdfc_ground_set_friction(self._h, float(mu)))

    def set_damping(self, damping):
        check(self._lib.sdfc_ground_set_damping(self._h, float(damping)))

    def force_sub(self, E, phi):
        """-> fm_sub (6,), n_active"""
        fm = np.empty(6)
        n = C.c_int32(0)
        check(self._lib.sdfc_ground_force(self._h, _dp(_colmajor44(E)), _dp(_f64(phi)), _dp(fm), C.byref(n)))
        self.touched |= n.value > 0
        return fm, n.value

    def force_deriv_sub(self, E, phi):
        """-> fm_sub (6,), Km_sub (6,6), Dm_sub (6,6), n_active"""
        fm, Km, Dm = np.empty(6), np.empty((6, 6)), np.empty((6, 6))
        n = C.c_int32(0)
        check(self._lib.sdfc_ground_force_deriv(self._h, _dp(_colmajor44(E)), _dp(_f64(phi)), _dp(fm), _dp(Km), _dp(Dm),
                                                C.byref(n)))
        self.touched |= n.value > 0
        return fm, np.ascontiguousarray(Km.T), np.ascontiguousarray(Dm.T), n.value

    def computeForce(self, fm, E, phi):
        """Accumulates into the system vector like ForceGroundContact.cpp:129."""
        f, n = self.force_sub(E, phi)
        fm[self.index_contact:self.index_contact + 6] += f
        return n

    def computeForceWithDerivative(self, fm, Km, Dm, E, phi):
        f, K, D, n = self.force_deriv_sub(E, phi)
        s = slice(self.index_contact, self.index_contact + 6)
        fm[s] += f
        Km[s, s] += K
        Dm[s, s] += D
        return n

    def close(self):
        if self._h:
            self._lib.sdfc_ground_destroy(self._h)
            self._h = C.c_void_p()


class ContactBatch:
    """E environments x P pairs sharing geometry (device-resident)."""

    def __init__(self, ctx, n_envs, n_bodies, pairs, body_a, body_b):
        self.ctx, self._lib = ctx, ctx._lib
        self.n_envs, self.n_bodies, self.n_pairs = int(n_envs), int(n_bodies), len(pairs)
        self.pairs = list(pairs)
        arr = (C.c_void_p * self.n_pairs)(*[p._h for p in pairs])
        a = np.ascontiguousarray(body_a, dtype=np.int32)
        b = np.ascontiguousarray(body_b, dtype=np.int32)
        self._h = C.c_void_p()
        check(self._lib.sdfc_batch_create(ctx._h, self.n_envs, self.n_bodies, self.n_pairs, arr,
                                          a.ctypes.data_as(_lib.c_int32_p), b.ctypes.data_as(_lib.c_int32_p),
                                          C.byref(self._h)))
        self.ndof_m = 6 * self.n_bodies

    @property
    def queries_per_eval(self):
        return int(self._lib.sdfc_batch_queries_per_eval(self._h))

    def set_grounds(self, grounds, bodies):
        """ForceGroundContact forces of every environment: grounds[g] acts on body bodies[g]."""
        self.grounds = list(grounds)
        arr = (C.c_void_p * len(grounds))(*[g._h for g in grounds])
        b = np.ascontiguousarray(bodies, dtype=np.int32)
        check(self._lib.sdfc_batch_set_grounds(self._h, len(grounds), arr, b.ctypes.data_as(_lib.c_int32_p)))
        self.n_grounds = len(grounds)

    def skin_stats(self):
        """{words_culled, margins_used_up, rebuilds, pairs_skipped} of the stepper's detection culling (zeros when off)."""
        s = (C.c_int64 * 4)()
        check(self._lib.sdfc_batch_get_skin_stats(self._h, s))
        return {"words_culled": int(s[0]), "margins_used_up": int(s[1]), "rebuilds": int(s[2]), "pairs_skipped": int(s[3])}

    def brick_stats(self):
        """{staged, too_large, outside}: detection tiles by how the TMA brick kernel served them (running totals)."""
        s = (C.c_int64 * 3)()
        check(self._lib.sdfc_batch_get_brick_stats(self._h, s))
        return {"staged": int(s[0]), "too_large": int(s[1]), "outside": int(s[2])}

    def ground_active(self):
        """(n_envs, n_grounds) contact counts of the last evaluation."""
        n = np.empty((self.n_envs, self.n_grounds), dtype=np.int32)
        check(self._lib.sdfc_batch_get_ground_active(self._h, n.ctypes.data_as(_lib.c_int32_p)))
        return n

    def eval(self, E, phi, deriv=True):
        """E: (n_envs, n_bodies, 4, 4); phi: (n_envs, n_bodies, 6). Host arrays in and out.
        -> fm (n_envs, ndof_m), Km, Dm (n_envs, ndof_m, ndof_m) or None, n_active (n_envs, n_pairs)."""
        Ec = _colmajor44(np.asarray(E).reshape(self.n_envs, self.n_bodies, 4, 4))
        ph = _f64(np.asarray(phi).reshape(self.n_envs, self.n_bodies, 6))
        nm = self.ndof_m
        fm = np.empty((self.n_envs, nm))
        Km = np.empty((self.n_envs, nm, nm)) if deriv else None
        Dm = np.empty((self.n_envs, nm, nm)) if deriv else None
        nact = np.empty((self.n_envs, self.n_pairs), dtype=np.int32)
        check(self._lib.sdfc_batch_eval(self._h, _dp(Ec), _dp(ph), int(bool(deriv)), _dp(fm),
                                        _dp(Km) if deriv else None, _dp(Dm) if deriv else None,
                                        nact.ctypes.data_as(_lib.c_int32_p)))
        if deriv:  # library writes column-major
            Km = np.ascontiguousarray(np.swapaxes(Km, 1, 2))
            Dm = np.ascontiguousarray(np.swapaxes(Dm, 1, 2))
        return fm, Km, Dm, nact

    def eval_into(self, E_col, phi, deriv, fm, Km, Dm, nact=None):
        """The host-buffer C-ABI call without any Python-side copies: E_col is
        (n_envs, n_bodies, 16) COLUMN-major poses, outputs are written in the library's
        layout (Km, Dm column-major per environment). Pass PinnedArray objects (or their `.a` views) for direct DMA."""
        def addr(x):   # PinnedArray: its stored address; numpy array: looked up
            return x.ptr if isinstance(x, PinnedArray) else (None if x is None else x.ctypes.data)
        check(self._lib.sdfc_batch_eval_addr(self._h, addr(E_col), addr(phi), int(bool(deriv)), addr(fm),
                                             addr(Km) if deriv else None, addr(Dm) if deriv else None, addr(nact)))

    def eval_device(self, E_ptr, phi_ptr, deriv, fm_ptr, Km_ptr, Dm_ptr, nact_ptr):
        """Raw device pointers (ints). E is column-major per pose. Asynchronous."""
        check(self._lib.sdfc_batch_eval_device(self._h, E_ptr, phi_ptr, int(bool(deriv)), fm_ptr, Km_ptr, Dm_ptr, nact_ptr))

    def active_mask(self, env, pair):
        """(n,) uint8 in ORIGINAL point order: the 1-bit active mask the last batched detection wrote for
        (env, pair) - collision_detection_general_SDF's d < 0 test (CollisionDetection.cpp:137-138)."""
        n = self.pairs[pair].contact_points.n
        out = np.empty(n, dtype=np.uint8)
        check(self._lib.sdfc_batch_get_active(self._h, int(env), int(pair), out.ctypes.data_as(_lib.c_uint8_p)))
        return out

    def detect_device(self, E_ptr, nact_ptr):
        check(self._lib.sdfc_batch_detect_device(self._h, E_ptr, nact_ptr))

    def close(self):
        if self._h:
            self._lib.sdfc_batch_destroy(self._h)
            self._h = C.c_void_p()


JOINT_FIXED, JOINT_TRANSLATIONAL, JOINT_FREE3D_EXP = 0, 1, 2
JOINT_TYPES = {"fixed": JOINT_FIXED, "translational": JOINT_TRANSLATIONAL, "free3d-exp": JOINT_FREE3D_EXP}
JOINT_NDOF = {JOINT_FIXED: 0, JOINT_TRANSLATIONAL: 3, JOINT_FREE3D_EXP: 6}


class BatchedSimulation:
    """Simulation::reset / set_state / forward / get_q ... (Simulation.cpp:73-122,894-993)
    for E environments at once, device-resident, on top of a ContactBatch.

    bodies: one dict per link in robot order with keys joint_type ("fixed" |
    "translational" | "free3d-exp"), E_pj_0 (4,4), E_ji (4,4), inertia (6,), mass.
    Reduced coordinates are ordered body by body (Robot.cpp:48-101)."""

    def __init__(self, batch, bodies, gravity, h):
        self.batch, self.ctx, self._lib = batch, batch.ctx, batch._lib
        if len(bodies) != batch.n_bodies:
            raise SdfContactError("[Error] BatchedSimulation: body count does not match the batch")
        arr = (_lib.BodyDesc * len(bodies))()
        self.joint_types = []
        for i, b in enumerate(bodies):
            jt = b["joint_type"]
            jt = JOINT_TYPES[jt] if isinstance(jt, str) else int(jt)
            self.joint_types.append(jt)
            arr[i].joint_type = jt
            arr[i].E_pj_0[:] = _colmajor44(b["E_pj_0"]).reshape(-1)
            arr[i].E_ji[:] = _colmajor44(b["E_ji"]).reshape(-1)
            arr[i].inertia[:] = _f64(b["inertia"]).reshape(6)
            arr[i].mass = float(b["mass"])
        g = _f64(gravity)
        self.h = float(h)
        self._h = C.c_void_p()
        check(self._lib.sdfc_sim_create(batch._h, arr, _dp(g), self.h, C.byref(self._h)))
        self.n_envs, self.n_bodies, self.n_pairs = batch.n_envs, batch.n_bodies, batch.n_pairs
        self.ndof_r = int(self._lib.sdfc_sim_ndof_r(self._h))
        self.dof_offset = np.cumsum([0] + [JOINT_NDOF[t] for t in self.joint_types])
        self._fext = np.zeros((self.n_envs, self.n_bodies, 6))

    # ---- state
    def set_state(self, q, qdot):
        q = _f64(np.broadcast_to(q, (self.n_envs, self.ndof_r)))
        qd = _f64(np.broadcast_to(qdot, (self.n_envs, self.ndof_r)))
        check(self._lib.sdfc_sim_set_state(self._h, _dp(q), _dp(qd)))

    def get_state(self):
        q = np.empty((self.n_envs, self.ndof_r))
        qd = np.empty((self.n_envs, self.ndof_r))
        check(self._lib.sdfc_sim_get_state(self._h, _dp(q), _dp(qd)))
        return q, qd

    def get_q(self):
        return self.get_state()[0]

    def get_qdot(self):
        return self.get_state()[1]

    def set_body_external_force(self, body, force):
        """force: (6,) for every environment or (n_envs, 6)."""
        self._fext[:, body, :] = np.broadcast_to(_f64(force), (self.n_envs, 6))
        check(self._lib.sdfc_sim_set_external_force(self._h, _dp(self._fext)))

    def reset(self, q=None, qdot=None):
        """Simulation::reset: initial state, cleared contact history and external forces."""
        self.set_state(np.zeros(self.ndof_r) if q is None else q, np.zeros(self.ndof_r) if qdot is None else qdot)
        check(self._lib.sdfc_sim_clear_contact_history(self._h))
        self._fext[:] = 0.0
        check(self._lib.sdfc_sim_set_external_force(self._h, _dp(self._fext)))

    def set_active(self, active=None):
        """Only environments with active[e] True take part in the following forward()
        calls (None = all); the others keep their state."""
        if active is None:
            check(self._lib.sdfc_sim_set_active(self._h, None))
            return
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(active, dtype=bool), (self.n_envs,)), dtype=np.uint8)
        check(self._lib.sdfc_sim_set_active(self._h, a.ctypes.data_as(_lib.c_uint8_p)))

    # ---- stepping
    def forward(self, n_steps, trajectory=False):
        """n BDF1 steps of every environment. trajectory=True -> (q, qdot) of shape
        (n_steps, n_envs, ndof_r) copied back once at the end."""
        if not trajectory:
            check(self._lib.sdfc_sim_forward(self._h, int(n_steps), None, None))
            return None
        Q = np.empty((n_steps, self.n_envs, self.ndof_r))
        QD = np.empty((n_steps, self.n_envs, self.ndof_r))
        check(self._lib.sdfc_sim_forward(self._h, int(n_steps), _dp(Q), _dp(QD)))
        return Q, QD

    def is_converged(self):
        c = np.empty(self.n_envs, dtype=np.uint8)
        check(self._lib.sdfc_sim_get_status(self._h, c.ctypes.data_as(_lib.c_uint8_p), None))
        return c.astype(bool)

    def get_contact_history(self):
        """(n_envs, n_pairs) bool: the pair recorded a contact since the last reset."""
        t = np.empty((self.n_envs, self.n_pairs), dtype=np.uint8)
        check(self._lib.sdfc_sim_get_status(self._h, None, t.ctypes.data_as(_lib.c_uint8_p)))
        return t.astype(bool)

    def get_ground_history(self):
        """(n_envs, n_grounds) bool: the body recorded "ground" since the last reset."""
        ng = getattr(self.batch, "n_grounds", 0)
        t = np.empty((self.n_envs, ng), dtype=np.uint8)
        check(self._lib.sdfc_sim_get_ground_history(self._h, t.ctypes.data_as(_lib.c_uint8_p)))
        return t.astype(bool)

    def get_body_E0i(self):
        """(n_envs, n_bodies, 4, 4) body poses at the current state."""
        E = np.empty((self.n_envs, self.n_bodies, 4, 4))
        check(self._lib.sdfc_sim_get_poses(self._h, _dp(E)))
        return np.ascontiguousarray(np.swapaxes(E, -1, -2))

    def stats(self):
        s = (C.c_int64 * 4)()
        check(self._lib.sdfc_sim_get_stats(self._h, s))
        return {"rounds": int(s[0]), "deriv_env_evals": int(s[1]), "value_env_evals": int(s[2]), "env_steps": int(s[3])}

    def close(self):
        if self._h:
            self._lib.sdfc_sim_destroy(self._h)
            self._h = C.c_void_p()


def build_simulation(ctx, scene, n_envs):
    """BatchedSimulation for `n_envs` copies of one constructed assembly.

    scene: {"bodies": [{points (n,3), grid {ni,nj,nk,min_box,max_box,dx,dy,dz,col_th,grid},
    joint_type, E_pj_0, E_ji, inertia, mass}, ...] in robot order, "forces": [{a, b, kn, kt,
    mu, damping}, ...] in Robot::_forces order, "h", "gravity"} - the state a constructed
    reference Simulation holds for this path (Simulation_Constructor.cpp:143-714)."""
    bodies, forces = scene["bodies"], scene["forces"]
    sdf_of, pts_of = {}, {}
    pairs = []
    for f in forces:
        a, b = f["a"], f["b"]
        if b not in sdf_of:
            sdf_of[b] = SdfGrid.from_dict(ctx, bodies[b]["grid"])
        if a not in pts_of:
            pts_of[a] = ContactPoints(ctx, bodies[a]["points"])
        pairs.append(ForceGeneralSDFContact(ctx, pts_of[a], sdf_of[b], f["kn"], f["kt"], f["mu"], f["damping"], 6 * a, 6 * b))
    batch = ContactBatch(ctx, n_envs, len(bodies), pairs, [f["a"] for f in forces], [f["b"] for f in forces])
    grounds = []
    for g in scene.get("grounds", []):
        a = g["body"]
        if a not in pts_of:
            pts_of[a] = ContactPoints(ctx, bodies[a]["points"])
        grounds.append(ForceGroundContact(ctx, pts_of[a], g["E_g"], g["kn"], g["kt"], g["mu"], g["damping"], 6 * a))
    if grounds:
        batch.set_grounds(grounds, [g["body"] for g in scene["grounds"]])
    sim = BatchedSimulation(batch, bodies, scene["gravity"], scene["h"])
    sim._keep = (batch, pairs, sdf_of, pts_of, grounds)
    return sim
