"""CPU: the stepper's warp-per-environment arithmetic (assemble-them-all_b200/csrc/sdfc_step_math.cuh),
compiled for the host (tests/host/step_harness.cpp), against the numpy oracle of the reference's
Robot::update / computeMatrices / BDF1 residual (oracle/sim_oracle.py, pinned to the reference by
tests/test_sim_oracle.py).

The header is a sequence of data-parallel statements (one output element per lane) separated by warp
barriers. On the host the elements of a statement run one after the other; running them FORWARDS and
BACKWARDS must give the same bits, which proves that no statement reads something the same statement
writes (i.e. that the barriers on the device are where they have to be)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from assemble_them_all_b200 import synth
from tests.helpers import ROOT

SRC = os.path.join(ROOT, "tests", "host", "step_harness.cpp")
HDR = os.path.join(ROOT, "assemble-them-all_b200", "csrc", "sdfc_step_math.cuh")
LIB = os.path.join(ROOT, "tests", "host", "libstep_harness.so")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def harness():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.step_eval.restype = C.c_int
    lib.step_eval.argtypes = [C.c_int, ip, dp, dp, dp, dp, C.c_double, dp, dp, dp, dp, dp, C.c_int, ip, ip, dp, dp, dp, C.c_int,
                              C.c_int, dp, dp, dp, dp, dp]
    lib.step_reparam.restype = None
    lib.step_reparam.argtypes = [dp, dp, C.c_int]
    return lib


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def _scene(rng, joints):
    from oracle import sim_oracle as S
    jt = {"fixed": S.FIXED, "translational": S.TRANSLATIONAL, "free3d-exp": S.FREE3D_EXP}
    bodies = []
    for j in joints:
        E_pj_0 = synth.make_E(synth.random_rotation(rng), rng.uniform(-1, 1, 3))
        E_ji = synth.make_E(synth.random_rotation(rng), rng.uniform(-0.5, 0.5, 3))
        bodies.append(S.BodyDesc(jt[j], E_pj_0, E_ji, rng.uniform(0.5, 3.0, 6), float(rng.uniform(1, 5))))
    return S, bodies


def _run(lib, S, bodies, h, grav, x, q0, qd0, fext, blocks, deriv, direction):
    nb = len(bodies)
    nr = sum(b.ndof for b in bodies)
    jt = np.array([b.joint_type for b in bodies], dtype=np.int32)
    Ep = np.ascontiguousarray(np.stack([b.E_pj_0.T for b in bodies]))   # column-major 4x4
    Ej = np.ascontiguousarray(np.stack([b.E_ji.T for b in bodies]))
    I6 = np.ascontiguousarray(np.stack([b.inertia for b in bodies]))
    mass = np.array([b.mass for b in bodies])
    pa = np.array([p[0] for p in blocks], dtype=np.int32)
    pb = np.array([p[1] for p in blocks], dtype=np.int32)
    fm = np.ascontiguousarray(np.stack([p[2] for p in blocks])) if blocks else np.zeros((0, 12))
    Km = np.ascontiguousarray(np.stack([p[3].T for p in blocks])) if blocks else np.zeros((0, 144))   # column-major
    Dm = np.ascontiguousarray(np.stack([p[4].T for p in blocks])) if blocks else np.zeros((0, 144))
    g, H, E, phi, dx = np.zeros(nr), np.zeros((nr, nr)), np.zeros((nb, 16)), np.zeros((nb, 6)), np.zeros(nr)
    got = lib.step_eval(nb, _i(jt), _d(Ep), _d(Ej), _d(I6), _d(mass), h, _d(np.ascontiguousarray(grav)), _d(x), _d(q0), _d(qd0),
                        _d(np.ascontiguousarray(fext)), len(blocks), _i(pa), _i(pb), _d(fm), _d(Km), _d(Dm), int(deriv), direction,
                        _d(g), _d(H), _d(E), _d(phi), _d(dx))
    assert got == nr
    return g, H, E.reshape(nb, 4, 4).transpose(0, 2, 1), phi, dx


def _oracle(S, bodies, h, grav, x, q0, qd0, fext, blocks, deriv):
    from oracle import sdfc_oracle as O
    sim = S.Sim(bodies, [], h, np.asarray(grav, dtype=np.float64))
    sim.f_ext = fext.copy()

    def contact_forces(E, phi, fm, Km, Dm, d):
        for a, b, f, K, D in blocks:
            if d:
                O.scatter(sim.ndof_m, 6 * a, 6 * b, f, K, D, fm, Km, Dm)
            else:
                fm[6 * a:6 * a + 6] += f[:6]
                fm[6 * b:6 * b + 6] += f[6:]
    sim.contact_forces = contact_forces
    E, phi, *_ = sim.kinematics(x, (x - q0) / h, False)
    if deriv:
        g, H = sim.eval_g_H(x, q0, qd0)
        return g, H, E, phi
    return sim.eval_g(x, q0, qd0), None, E, phi


def _rel(a, b):
    s = np.max(np.abs(a))
    return float(np.max(np.abs(a - b)) / s) if s > 0 else float(np.max(np.abs(b)))


CONFIGS = [("free3d-exp", "free3d-exp"), ("free3d-exp", "fixed"), ("fixed", "free3d-exp"), ("translational", "fixed"),
           ("translational", "free3d-exp"), ("free3d-exp", "fixed", "fixed")]


@pytest.mark.parametrize("joints", CONFIGS)
def test_residual_jacobian_and_poses_match_oracle(harness, joints):
    from oracle import sdfc_oracle as O
    O.build()
    rng = np.random.default_rng(100 + CONFIGS.index(joints))
    S, bodies = _scene(rng, joints)
    nr = sum(b.ndof for b in bodies)
    nb = len(bodies)
    h = float(np.float32(1e-3))
    worst = 0.0
    for trial in range(12):
        grav = rng.normal(size=3) * (9.8 if trial % 2 else 1e-12)
        q0 = rng.normal(size=nr) * (0.3 if trial < 6 else 2.0)
        qd0 = rng.normal(size=nr)
        x = q0 + h * qd0 + rng.normal(size=nr) * 1e-4
        if trial == 3:   # exact zero rotation: the closed-form branch of JointSphericalExp::update
            off = 0
            for b in bodies:
                if b.ndof == 6:
                    x[off + 3:off + 6] = 0.0
                off += b.ndof
        fext = rng.normal(size=(nb, 6)) * 50
        blocks = []
        if nb >= 2:
            for a, b in ((0, 1), (1, 0)):
                blocks.append((a, b, rng.normal(size=12) * 1e3, rng.normal(size=(12, 12)) * 1e5, rng.normal(size=(12, 12)) * 10))
        for deriv in (True, False):
            g1, H1, E1, phi1, dx1 = _run(harness, S, bodies, h, grav, x, q0, qd0, fext, blocks, deriv, +1)
            g2, H2, E2, phi2, dx2 = _run(harness, S, bodies, h, grav, x, q0, qd0, fext, blocks, deriv, -1)
            for a, b in ((g1, g2), (H1, H2), (E1, E2), (phi1, phi2), (dx1, dx2)):
                assert np.array_equal(a, b), "a statement depends on the order of its own elements (missing barrier)"
            assert not np.isnan(g1).any() and not np.isnan(E1).any() and not np.isnan(phi1).any()
            gr, Hr, Er, phir = _oracle(S, bodies, h, grav, x, q0, qd0, fext, blocks, deriv)
            errs = [_rel(gr, g1), _rel(Er, E1), _rel(phir, phi1)]
            if deriv:
                assert not np.isnan(H1).any() and not np.isnan(dx1).any()
                errs.append(_rel(Hr, H1))
                ref_dx = np.linalg.solve(Hr, -gr)
                assert _rel(ref_dx, dx1) <= 1e-9 * max(1.0, np.linalg.cond(Hr))
            worst = max(worst, *errs)
            assert max(errs) <= 1e-11, (joints, trial, deriv, errs)
    print(f"{joints}: worst relative difference {worst:.2e}")


def test_reparam_matches_oracle(harness):
    from oracle import sim_oracle as S
    rng = np.random.default_rng(5)
    bd = S.BodyDesc(S.FREE3D_EXP, np.eye(4), np.eye(4), np.ones(6), 1.0)
    for scale in (1.0, 5.0, 7.5, 12.0, 30.0):
        axis = rng.normal(size=3)
        r = axis / np.linalg.norm(axis) * scale
        rd = rng.normal(size=3)
        q, qd = np.concatenate([np.zeros(3), r]), np.concatenate([np.zeros(3), rd])
        qr, qdr = S.reparam(bd, q, qd)
        for direction in (+1, -1):
            r1, rd1 = r.copy(), rd.copy()
            harness.step_reparam(_d(r1), _d(rd1), direction)
            assert _rel(qr[3:], r1) <= 1e-12 and _rel(qdr[3:], rd1) <= 1e-9, (scale, direction)
        assert (np.linalg.norm(r) > 1.5 * np.pi) == (not np.array_equal(qr[3:], r))
