"""CPU restatement (numpy) of the reference's BDF1 step around the contact path.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg, never by the product package.

Restates, for "flat" robots (every link is a root link with its own joint, which
is what examples/run_joint_plan.py:140-178 and test_joint_sim.py:24-63 emit), the
caller side of the hot path. Paths are relative to
/root/reference/simulation/redmax:

  joint kinematics   Joint/Joint.cpp:95-141, Joint/JointFixed.cpp:6-13,
                     Joint/JointTranslational.cpp:9-40, Joint/JointFree3DExp.cpp:13-110,
                     Joint/JointSphericalExp.cpp:13-183 (update), :237-264 (reparam)
  body kinematics    Body/Body.cpp:138-181
  body forces        Body/Body.cpp:250-330 (Coriolis, gravity, external)
  Jacobians          Robot.cpp:656-744 (root joints only: no ancestor loop)
  assembly           Simulation.cpp:379-503 (computeMatrices, both overloads)
  BDF1 residual      Simulation.cpp:1081-1112
  Newton             Simulation.cpp:999-1079 (line search rules, strike logic)
  step               Simulation.cpp:914-966, 1248-1270
  contact forces     oracle/sdfc_oracle.c (pair_force) in force order, Robot.cpp:524-547

Parity status: pinned against per-step q / qdot of the unmodified reference
(tests/golden/traj_small.npz, produced by oracle/make_golden.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import sdfc_oracle as O

FIXED, TRANSLATIONAL, FREE3D_EXP = 0, 1, 2
NDOF = {FIXED: 0, TRANSLATIONAL: 3, FREE3D_EXP: 6}
MATH_EPS = 1e-9  # math::eps, Utils.h:14
PI = float(np.arccos(-1.0))


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def unskew(a):
    return np.array([a[2, 1], a[0, 2], a[1, 0]])


def Einv(E):
    out = np.eye(4)
    out[:3, :3] = E[:3, :3].T
    out[:3, 3] = -E[:3, :3].T @ E[:3, 3]
    return out


def Ad(E):  # Utils.h:88-97
    R = E[:3, :3]
    A = np.zeros((6, 6))
    A[:3, :3] = R
    A[3:, :3] = skew(E[:3, 3]) @ R
    A[3:, 3:] = R
    return A


def ad(phi):  # Utils.h:99-109
    a = np.zeros((6, 6))
    a[:3, :3] = skew(phi[:3])
    a[3:, :3] = skew(phi[3:])
    a[3:, 3:] = skew(phi[:3])
    return a


def exp3(w):  # Utils.h:111-127
    wlen = np.linalg.norm(w)
    R = np.eye(3)
    if wlen > 1e-12:
        wx, wy, wz = w / wlen
        c, s = np.cos(wlen), np.sin(wlen)
        c1 = 1.0 - c
        R = np.array([[c + wx * wx * c1, -wz * s + wx * wy * c1, wy * s + wx * wz * c1],
                      [wz * s + wx * wy * c1, c + wy * wy * c1, -wx * s + wy * wz * c1],
                      [-wy * s + wx * wz * c1, wx * s + wy * wz * c1, c + wz * wz * c1]])
    return R


def spherical_exp(q, qdot, deriv=True):
    """JointSphericalExp::update (JointSphericalExp.cpp:13-183).
    -> R, Rdot, T (=S_j top 3x3), Tdot, dR_dq[k], dRdot_dq[k], dT_dq[k], dTdot_dq[k]."""
    I = np.eye(3)
    e = [I[:, j] for j in range(3)]
    z3 = [np.zeros((3, 3)) for _ in range(3)]
    if np.linalg.norm(q) < MATH_EPS:
        R = I.copy()
        dR = [skew(e[j]) for j in range(3)]
        Rdot = sum(skew(e[j]) * qdot[j] for j in range(3))
        T = I.copy()
        Tdot = np.zeros((3, 3))
        for j in range(3):
            Tdot[:, j] = unskew(Rdot.T @ skew(e[j]))
        dT = [np.zeros((3, 3)) for _ in range(3)]
        for j in range(3):
            for k in range(3):
                dT[k][:, j] = unskew(skew(e[k]).T @ skew(e[j]))
        return R, Rdot, T, Tdot, dR, z3, dT, [np.zeros((3, 3)) for _ in range(3)]
    R = exp3(q)
    q2 = float(q @ q)
    d = 1.0 / q2
    A, B, Cc = [None] * 3, [None] * 3, [None] * 3
    for j in range(3):
        Cc[j] = skew(skew(q) @ ((I - R) @ e[j]))
        B[j] = q[j] * skew(q)
        A[j] = (B[j] + Cc[j]) * d
    dR = [A[j] @ R for j in range(3)]
    Rdot = sum(dR[j] * qdot[j] for j in range(3))
    ddot = -2.0 * float(q @ qdot) / (q2 * q2)
    Adot, Bdot, Cdot = [None] * 3, [None] * 3, [None] * 3
    for j in range(3):
        Cdot[j] = skew(skew(qdot) @ ((I - R) @ e[j]) - skew(q) @ (Rdot @ e[j]))
        Bdot[j] = qdot[j] * skew(q) + q[j] * skew(qdot)
        Adot[j] = (Bdot[j] + Cdot[j]) * d + (B[j] + Cc[j]) * ddot
    T = np.zeros((3, 3))
    Tdot = np.zeros((3, 3))
    for j in range(3):
        T[:, j] = unskew(R.T @ dR[j])
        Tdot[:, j] = unskew(Rdot.T @ A[j] @ R + R.T @ Adot[j] @ R + R.T @ A[j] @ Rdot)
    if not deriv:
        return R, Rdot, T, Tdot, dR, None, None, None
    dd_dq = np.array([-2.0 * q[k] / (q2 * q2) for k in range(3)])
    dA = [[None] * 3 for _ in range(3)]   # dA[j][k]
    dB = [[None] * 3 for _ in range(3)]
    dC = [[None] * 3 for _ in range(3)]
    for k in range(3):
        for j in range(3):
            dB[j][k] = q[j] * skew(e[k])
            if j == k:
                dB[j][k] = dB[j][k] + skew(q)
            dC[j][k] = skew(skew(e[k]) @ ((I - R) @ e[j])) + skew(-skew(q) @ (dR[k] @ e[j]))
            dA[j][k] = (dB[j][k] + dC[j][k]) * d + (B[j] + Cc[j]) * dd_dq[k]
    dRdot = [np.zeros((3, 3)) for _ in range(3)]
    for k in range(3):
        for j in range(3):
            d2R = dA[j][k] @ R + A[j] @ dR[k]
            dRdot[k] = dRdot[k] + d2R * qdot[j]
    dddot_dq = np.array([(-2.0 * qdot[k] * q2 + 8.0 * float(q @ qdot) * q[k]) / q2 ** 3 for k in range(3)])
    dAdot = [[None] * 3 for _ in range(3)]
    for k in range(3):
        for j in range(3):
            dBdot = qdot[j] * skew(e[k])
            if j == k:
                dBdot = dBdot + skew(qdot)
            dCdot = skew(-skew(qdot) @ (dR[k] @ e[j]) - skew(e[k]) @ (Rdot @ e[j]) - skew(q) @ (dRdot[k] @ e[j]))
            dAdot[j][k] = ((dBdot + dCdot) * d + (Bdot[j] + Cdot[j]) * dd_dq[k] + (dB[j][k] + dC[j][k]) * ddot
                           + (B[j] + Cc[j]) * dddot_dq[k])
    dT = [np.zeros((3, 3)) for _ in range(3)]
    dTdot = [np.zeros((3, 3)) for _ in range(3)]
    for j in range(3):
        for k in range(3):
            d2R = dA[j][k] @ R + A[j] @ dR[k]
            dT[k][:, j] = unskew(dR[k].T @ dR[j] + R.T @ d2R)
            dTdot[k][:, j] = unskew(
                dRdot[k].T @ A[j] @ R + Rdot.T @ dA[j][k] @ R + Rdot.T @ A[j] @ dR[k]
                + dR[k].T @ Adot[j] @ R + R.T @ dAdot[j][k] @ R + R.T @ Adot[j] @ dR[k]
                + dR[k].T @ A[j] @ Rdot + R.T @ dA[j][k] @ Rdot + R.T @ A[j] @ dRdot[k])
    return R, Rdot, T, Tdot, dR, dRdot, dT, dTdot


@dataclass
class BodyDesc:
    """What the stepper needs from one <link>: joint type + Joint::_E_pj_0, Body::_E_ji,
    Body::_Inertia (6), Body::_mass."""
    joint_type: int
    E_pj_0: np.ndarray
    E_ji: np.ndarray
    inertia: np.ndarray
    mass: float

    @property
    def ndof(self):
        return NDOF[self.joint_type]


@dataclass
class PairDesc:
    """One ForceGeneralSDFContact in force order: contact body a, SDF body b."""
    a: int
    b: int
    grid: O.Grid
    points: np.ndarray
    kn: float
    kt: float
    mu: float = 0.0
    damping: float = 0.0


@dataclass
class GroundDesc:
    """One ForceGroundContact (Force/ForceGroundContact.cpp:9-20): contact body + ground frame."""
    body: int
    points: np.ndarray
    E_g: np.ndarray
    kn: float
    kt: float
    mu: float = 0.0
    damping: float = 0.0


CONST_EPS = 1e-8  # constants::eps, Utils.h:9


def _gamma(x):  # math::gamma, Utils.h:196-201
    return np.hstack([-skew(x), np.eye(3)])


def ground_detect(E_g, E, points):
    """collision_detection_ground_object (CollisionDetection.cpp:37-68), mesh branch:
    d = ng . (xw - xg) per contact point, active iff d <= 0. -> (d, active mask)."""
    xg, ng = E_g[:3, 3], E_g[:3, 2]
    R, p = E[:3, :3], E[:3, 3]
    X = np.asarray(points)
    # xw = E_0i * [xi; 1]: ((R0 x + R1 y) + R2 z) + p, no contraction
    xw = ((R[:, 0][None, :] * X[:, 0:1] + R[:, 1][None, :] * X[:, 1:2]) + R[:, 2][None, :] * X[:, 2:3]) + p[None, :]
    v = xw - xg[None, :]
    d = (ng[0] * v[:, 0] + ng[1] * v[:, 1]) + ng[2] * v[:, 2]
    return d, d <= 0.0


def ground_force(g: GroundDesc, E, phi, deriv):
    """ForceGroundContact::computeForce / computeForceWithDerivative
    (Force/ForceGroundContact.cpp:106-242). -> fm[6], Km[6,6], Dm[6,6], n_active, n_static."""
    xg, ng = g.E_g[:3, 3], g.E_g[:3, 2]
    N = np.outer(ng, ng)
    I3 = np.eye(3)
    T = I3 - N
    R, p = E[:3, :3], E[:3, 3]
    eb = [skew(I3[:, k]) for k in range(3)]
    RNR = R.T @ N @ R
    pxg = skew(R.T @ N @ (p - xg))
    fm = np.zeros(6)
    Km = np.zeros((6, 6))
    Dm = np.zeros((6, 6))
    d_all, act = ground_detect(g.E_g, E, g.points)
    n_static = 0
    for i in np.nonzero(act)[0]:
        xi = g.points[i]
        d = d_all[i]
        G = _gamma(xi)
        J = R @ G
        Gphi = G @ phi
        vwi = R @ Gphi
        fc = -g.kn * ng * d - g.damping * (N @ vwi)                       # :127 / :173
        fm += J.T @ fc
        if deriv:
            RNRxi = RNR @ xi
            RNRGphi = RNR @ Gphi
            tmp1 = np.column_stack([-e @ RNRxi for e in eb]) - RNR @ skew(xi) + pxg       # :176-178
            tmp2 = np.column_stack([-e @ RNRGphi for e in eb]) - RNR @ skew(Gphi)         # :180-182
            Km -= g.kn * G.T @ np.hstack([tmp1, RNR]) + g.damping * G.T @ np.hstack([tmp2, np.zeros((3, 3))])
            Dm -= g.damping * G.T @ RNR @ G
        if g.mu < CONST_EPS:                                              # :132 / :191
            continue
        a = T @ vwi
        an = np.linalg.norm(a)
        if g.mu * abs(g.kn * d) >= g.kt * an - CONST_EPS:                 # :137 / :197
            n_static += 1
            fm += J.T @ (-g.kt * a)
            if deriv:
                B = R.T @ T @ R
                tmp3 = np.hstack([np.column_stack([(B @ e - e @ B) @ Gphi for e in eb]), np.zeros((3, 3))])
                Dm += -g.kt * J.T @ T @ J
                Km += -g.kt * G.T @ tmp3
        else:
            mukn = g.mu * g.kn
            t = a / an
            fm += J.T @ (mukn * d * t)
            if deriv:
                A = (I3 - np.outer(t, t)) / an
                Dm += mukn * J.T @ (d * A) @ T @ J
                Rt = R.T @ t
                K1 = -d * np.hstack([np.column_stack([e @ Rt for e in eb]), np.zeros((3, 3))])
                K2 = np.outer(R.T @ t, ng) @ J
                K3 = -d * R.T @ A @ T @ R @ np.hstack([skew(Gphi), np.zeros((3, 3))])
                Km += mukn * G.T @ (K1 + K2 + K3)
    return fm, Km, Dm, int(act.sum()), n_static


def joint_kinematics(bd: BodyDesc, q, qdot, deriv):
    """-> Q (4x4), S_j, S_j_dot (6 x ndof), dSj_dq[k], dSjdot_dq[k]."""
    n = bd.ndof
    Q = np.eye(4)
    S = np.zeros((6, n))
    Sd = np.zeros((6, n))
    dS = [np.zeros((6, n)) for _ in range(n)]
    dSd = [np.zeros((6, n)) for _ in range(n)]
    if bd.joint_type == TRANSLATIONAL:
        Q[:3, 3] = q
        S[3:, :] = np.eye(3)
    elif bd.joint_type == FREE3D_EXP:
        p, r, rdot = q[:3], q[3:], qdot[3:]
        R, Rdot, T, Tdot, dR, dRdot, dT, dTdot = spherical_exp(r, rdot, deriv)
        Q[:3, :3] = R
        Q[:3, 3] = p
        S[:3, 3:] = T
        S[3:, :3] = R.T
        Sd[:3, 3:] = Tdot
        Sd[3:, :3] = Rdot.T
        if deriv:
            for i in range(3, 6):
                dS[i][:3, 3:] = dT[i - 3]
                dS[i][3:, :3] = dR[i - 3].T
                dSd[i][:3, 3:] = dTdot[i - 3]
                dSd[i][3:, :3] = dRdot[i - 3].T
    return Q, S, Sd, dS, dSd


def reparam(bd: BodyDesc, q, qdot):
    """JointFree3DExp::reparam -> JointSphericalExp::reparam (JointSphericalExp.cpp:237-264)."""
    if bd.joint_type != FREE3D_EXP:
        return q, qdot
    r, rdot = q[3:].copy(), qdot[3:].copy()
    qn = np.linalg.norm(r)
    if qn > 1.5 * PI:
        _, _, S0, *_ = spherical_exp(r, rdot, False)  # _S_j of the last update (same q, qdot)
        while qn > 1.5 * PI:
            r = (1.0 - 2.0 * PI / qn) * r
            qn = np.linalg.norm(r)
        _, _, S1, *_ = spherical_exp(r, rdot, False)
        rdot = np.linalg.solve(S1, S0 @ rdot)
        q, qdot = q.copy(), qdot.copy()
        q[3:], qdot[3:] = r, rdot
    return q, qdot


@dataclass
class Sim:
    bodies: list
    pairs: list
    h: float
    gravity: np.ndarray
    q: np.ndarray = None
    qdot: np.ndarray = None
    f_ext: np.ndarray = None          # (n_bodies, 6)
    converged: bool = True
    grounds: list = field(default_factory=list)   # GroundDesc, evaluated before the SDF pairs (XML order)
    touched: np.ndarray = None        # per pair, contact history (CollisionDetection.cpp:143-146)
    touched_ground: np.ndarray = None
    stats: dict = field(default_factory=lambda: {"deriv_evals": 0, "value_evals": 0, "newton_iters": 0})

    def __post_init__(self):
        self.off = np.cumsum([0] + [b.ndof for b in self.bodies])
        self.ndof_r = int(self.off[-1])
        self.ndof_m = 6 * len(self.bodies)
        if self.q is None:
            self.q = np.zeros(self.ndof_r)
        if self.qdot is None:
            self.qdot = np.zeros(self.ndof_r)
        if self.f_ext is None:
            self.f_ext = np.zeros((len(self.bodies), 6))
        self.touched = np.zeros(len(self.pairs), dtype=bool)
        self.touched_ground = np.zeros(len(self.grounds), dtype=bool)
        self.A_ij = [Ad(Einv(b.E_ji)) for b in self.bodies]

    # ---- kinematics: Robot::update for root joints
    def kinematics(self, q, qdot, deriv):
        nb = len(self.bodies)
        E = np.zeros((nb, 4, 4))
        phi = np.zeros((nb, 6))
        J = np.zeros((self.ndof_m, self.ndof_r))
        Jd = np.zeros((self.ndof_m, self.ndof_r))
        dJ = [np.zeros((self.ndof_m, self.ndof_r)) for _ in range(self.ndof_r)]
        dJd = [np.zeros((self.ndof_m, self.ndof_r)) for _ in range(self.ndof_r)]
        for i, bd in enumerate(self.bodies):
            o, n = self.off[i], bd.ndof
            Q, S, Sd, dS, dSd = joint_kinematics(bd, q[o:o + n], qdot[o:o + n], deriv)
            E_0j = bd.E_pj_0 @ Q                      # Joint.cpp:101,112-113
            E[i] = E_0j @ bd.E_ji                     # Body.cpp:140
            if n:
                phi_j = S @ qdot[o:o + n]             # Joint.cpp:119-121
                phi[i] = self.A_ij[i] @ phi_j         # Body.cpp:156
                J[6 * i:6 * i + 6, o:o + n] = self.A_ij[i] @ S      # Robot.cpp:664-665
                Jd[6 * i:6 * i + 6, o:o + n] = self.A_ij[i] @ Sd
                if deriv:
                    for k in range(n):
                        dJ[o + k][6 * i:6 * i + 6, o:o + n] = self.A_ij[i] @ dS[k]
                        dJd[o + k][6 * i:6 * i + 6, o:o + n] = self.A_ij[i] @ dSd[k]
        return E, phi, J, Jd, dJ, dJd

    # ---- Body::computeMaximalForce[WithDerivative] (Body.cpp:250-330)
    def body_forces(self, E, phi, deriv):
        nm = self.ndof_m
        fm = np.zeros(nm)
        Km = np.zeros((nm, nm)) if deriv else None
        Dm = np.zeros((nm, nm)) if deriv else None
        for i, bd in enumerate(self.bodies):
            R = E[i][:3, :3]
            I6 = bd.inertia
            adT = ad(phi[i]).T
            f_cor = adT @ (I6 * phi[i])
            f_g = np.zeros(6)
            f_g[3:] = R.T @ (bd.mass * self.gravity)
            f_e = np.zeros(6)
            f_e[:3] = I6[:3] * (R.T @ self.f_ext[i][:3])
            f_e[3:] = I6[3:] * (R.T @ self.f_ext[i][3:])
            s = slice(6 * i, 6 * i + 6)
            fm[s] += f_cor + f_g + f_e
            if deriv:
                Km[6 * i + 3:6 * i + 6, 6 * i:6 * i + 3] += skew(f_g[3:])
                df = np.zeros((6, 3))
                df[:3] = skew(R.T @ self.f_ext[i][:3])
                df[3:] = skew(R.T @ self.f_ext[i][3:])
                df = df * I6[:, None]
                Km[s, 6 * i:6 * i + 3] += df
                for k in range(6):
                    Dm[s, 6 * i + k] += adT[:, k] * I6[k]
                Iw = I6[:3] * phi[i][:3]
                mv = I6[3:] * phi[i][3:]
                for k in range(3):
                    ek = skew(np.eye(3)[:, k])
                    Dm[6 * i:6 * i + 3, 6 * i + k] -= ek @ Iw
                    Dm[6 * i:6 * i + 3, 6 * i + 3 + k] -= ek @ mv
                    Dm[6 * i + 3:6 * i + 6, 6 * i + k] -= ek @ mv
        return fm, Km, Dm

    def contact_forces(self, E, phi, fm, Km, Dm, deriv):
        for gi, g in enumerate(self.grounds):
            f, K, D, cnt, _ = ground_force(g, E[g.body], phi[g.body], deriv)
            s = slice(6 * g.body, 6 * g.body + 6)
            fm[s] += f
            if cnt > 0:
                self.touched_ground[gi] = True
            if deriv:
                Km[s, s] += K
                Dm[s, s] += D
        for pi, p in enumerate(self.pairs):
            f, K, D, cnt = O.pair_force(p.grid, E[p.a], phi[p.a], E[p.b], phi[p.b], p.points, p.kn, p.kt, p.mu,
                                        p.damping, deriv)
            if cnt[0] > 0:
                self.touched[pi] = True
            if deriv:
                O.scatter(self.ndof_m, 6 * p.a, 6 * p.b, f, K, D, fm, Km, Dm)
            else:
                fm[6 * p.a:6 * p.a + 6] += f[:6]
                fm[6 * p.b:6 * p.b + 6] += f[6:]

    def mass_diag(self):
        return np.concatenate([b.inertia for b in self.bodies])

    # ---- Simulation::computeMatrices (Simulation.cpp:379-503)
    def compute_matrices(self, q, qdot, deriv):
        E, phi, J, Jd, dJ, dJd = self.kinematics(q, qdot, deriv)
        Mm = self.mass_diag()
        fm, Km, Dm = self.body_forces(E, phi, deriv)
        self.contact_forces(E, phi, fm, Km, Dm, deriv)
        self.stats["deriv_evals" if deriv else "value_evals"] += 1
        MmJ = J * Mm[:, None]
        MmJd = Jd * Mm[:, None]
        M = J.T @ MmJ
        if not deriv:
            f = J.T @ (fm - MmJd @ qdot)
            return M, f
        nr = self.ndof_r
        JTMm = MmJ.T
        JTMmJd = JTMm @ Jd
        f = J.T @ fm + (-JTMmJd @ qdot)
        dM = []
        for k in range(nr):
            tmp = dJ[k].T @ MmJ
            dM.append(tmp + tmp.T)
        dJq = np.zeros((self.ndof_m, nr))
        for k in range(nr):
            dJq[:, k] = dJ[k] @ qdot
        Dqvv = -JTMmJd - JTMm @ dJq
        MmJdq = MmJd @ qdot
        Kqvv = np.zeros((nr, nr))
        for k in range(nr):
            Kqvv[:, k] = -dJ[k].T @ MmJdq - JTMm @ (dJd[k] @ qdot)
        JTDm = J.T @ Dm
        K = Kqvv + J.T @ Km @ J + JTDm @ dJq
        for k in range(nr):
            K[:, k] += dJ[k].T @ fm
        D = Dqvv + JTDm @ J
        return M, f, dM, K, D

    # ---- BDF1 residual (Simulation.cpp:1081-1112)
    def eval_g(self, q1, q0, qdot0):
        h = self.h
        qdot1 = (q1 - q0) / h
        M, f = self.compute_matrices(q1, qdot1, False)
        return M @ (q1 - q0 - h * qdot0) - h * h * f

    def eval_g_H(self, q1, q0, qdot0):
        h = self.h
        qdot1 = (q1 - q0) / h
        M, f, dM, K, D = self.compute_matrices(q1, qdot1, True)
        dq = q1 - q0 - h * qdot0
        g = M @ dq - h * h * f
        H = M - h * h * K - h * D
        for k in range(self.ndof_r):
            H[:, k] += dM[k] @ dq
        return g, H

    # ---- Simulation::newton (Simulation.cpp:999-1079)
    def newton(self, x, q0, qdot0):
        if self.ndof_r == 0:
            return x
        tol = 1e-8
        max_iter = max(20 * self.ndof_r, 100)
        success = False
        fail_strike = 0
        for _ in range(max_iter):
            g, H = self.eval_g_H(x, q0, qdot0)
            self.stats["newton_iters"] += 1
            dx = np.linalg.solve(H, -g)   # PartialPivLU
            gnorm = np.linalg.norm(g)
            alpha = 1.0
            ok = False
            g_new = None
            for _trial in range(20):
                g_new = self.eval_g(x + alpha * dx, q0, qdot0)
                if np.linalg.norm(g_new) < gnorm:
                    ok = True
                    break
                alpha *= 0.5
            if ok:
                fail_strike = 0
            else:
                fail_strike += 1
                if fail_strike >= 3:
                    break
            x = x + alpha * dx
            if np.linalg.norm(g_new) < tol:
                success = True
                break
        self.converged = success
        return x

    # ---- Simulation::forward, one BDF1 step (Simulation.cpp:914-966, 1248-1257)
    def step(self):
        q0, qdot0 = self.q.copy(), self.qdot.copy()
        x = self.newton(q0 + self.h * qdot0, q0, qdot0)
        qdot1 = (x - q0) / self.h
        for i, bd in enumerate(self.bodies):
            o, n = self.off[i], bd.ndof
            if n:
                x[o:o + n], qdot1[o:o + n] = reparam(bd, x[o:o + n], qdot1[o:o + n])
        self.q, self.qdot = x, qdot1

    def body_poses(self):
        E, phi, *_ = self.kinematics(self.q, self.qdot, False)
        return E, phi
