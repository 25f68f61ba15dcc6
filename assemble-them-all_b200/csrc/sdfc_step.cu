// sdfc_step.cu — device-resident batched BDF1 stepper: the
// caller side of the contact path for E environments at once, so that
// Simulation::forward(n) never returns to the host between substeps.
//
// One evaluation "round" for all environments that still have work:
//   k_sim_pre    candidate state -> body poses E_0i and twists phi (joint + body
//                kinematics), per-environment evaluation mode (value / derivative)
//   batch_run    detection + contact force (+ Jacobians) kernels of sdfc_kernels.cu
//   k_sim_post   body forces, reduced assembly, BDF1 residual g (and H), LU solve,
//                line-search / Newton / step state machine
// Every environment advances through its own Newton iterations and line-search
// trials; the host only polls one counter (environments not yet finished).
//
// Restated from (paths relative to simulation/redmax/ of the reference):
//   Joint/Joint.cpp:95-141, JointFixed.cpp:6-13, JointTranslational.cpp:9-40,
//   JointFree3DExp.cpp:13-110, JointSphericalExp.cpp:13-183,237-264,
//   Body/Body.cpp:138-181,250-330, Robot.cpp:656-744 (root joints),
//   Simulation.cpp:379-503 (computeMatrices), :999-1079 (newton),
//   :1081-1112 (BDF1 residual), :914-966,1248-1257 (forward / integration_BDF1).
// One thread per environment: everything here is O(100) - O(10^4) flops per
// evaluation against O(10^7) in the contact kernels.
#include "sdfc_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#ifndef SDFC_SKIN_DEFAULT
#define SDFC_SKIN_DEFAULT 1
#endif

namespace sdfc {

constexpr int kCapR = 12;    // reduced dofs per environment (largest instantiation)
constexpr int kCapDyn = 2;   // non-fixed bodies per environment
// The per-environment kernels are instantiated for <max reduced dofs, max non-fixed bodies,
// any free3d-exp joint>: the planners' translational case (3, 1, false) keeps its state in
// registers instead of the ~20 KB of local memory the general (12, 2, true) case needs.
constexpr double kMathEps = 1e-9;  // math::eps, Utils.h:14
constexpr double kNewtonTol = 1e-8;  // Simulation.cpp:1008
constexpr int kMaxLS = 20;           // Simulation.cpp:1012
constexpr int kMaxStrike = 3;        // Simulation.cpp:1014

struct BodyDev {
    int type, dof_off, ndof, dyn;  // dyn: slot among non-fixed bodies, or -1
    double Rp[9], pp[3];           // Joint::_E_pj_0 (row-major rotation)
    double Rji[9], pji[3];         // Body::_E_ji
    double Aij[36];                // Ad(E_ij), row-major
    double I6[6], mass;
};

struct SimDev {
    const BodyDev* bodies;
    const int *pair_a, *pair_b;
    int n_bodies, n_pairs, n_envs, nr, n_dyn, max_iter;
    double h, grav[3];
    double *q, *qdot, *q0, *qdot0, *x, *dx;  // [n_envs][nr]
    double *gnorm, *alpha;                   // [n_envs]
    int *phase, *trial, *iter, *strike, *steps_left, *step_idx;
    uint8_t *converged, *touched;
    const uint8_t* active;                   // [n_envs] or nullptr (all environments step)
    const double* fext;                      // [n_envs][n_bodies][6]
    double *E, *phi;                         // contact-path inputs
    int* env_mode;
    const double *fm_sub, *Km_sub, *Dm_sub;  // contact-path outputs per (env, pair)
    const int32_t* n_active;
    int n_grounds;                           // ForceGroundContact forces (0 = none)
    const int* ground_body;                  // [n_grounds]
    const double* g_out;                     // [n_envs][n_grounds][78]: fm 6, Km 36, Dm 36 (column-major)
    const int32_t* g_n_active;               // [n_envs][n_grounds]
    uint8_t* touched_ground;                 // [n_envs][n_grounds]
    double *q_traj, *qdot_traj;              // [steps][n_envs][nr] or nullptr
    int* n_unfinished;
    unsigned long long* counters;            // deriv evals, value evals, steps
    int env0, env_count;                     // environment range of this launch (a stepper lane)
};

enum { PH_DERIV = 0, PH_LS = 1, PH_DONE = 2 };

// ---------------------------------------------------------------- 3x3 helpers (row-major)
__device__ __forceinline__ void skew(const double* v, double* S) {
    S[0] = 0; S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2]; S[4] = 0; S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
__device__ __forceinline__ void unskew(const double* A, double* v) { v[0] = A[7]; v[1] = A[2]; v[2] = A[3]; }
__device__ __noinline__ void mul33(const double* A, const double* B, double* C) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}
__device__ __noinline__ void tmul33(const double* A, const double* B, double* C) {  // A^T B
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
}
__device__ __forceinline__ void mulv3(const double* A, const double* v, double* o) {
    for (int r = 0; r < 3; ++r) o[r] = A[r * 3] * v[0] + A[r * 3 + 1] * v[1] + A[r * 3 + 2] * v[2];
}
__device__ __forceinline__ void tmulv3(const double* A, const double* v, double* o) {
    for (int r = 0; r < 3; ++r) o[r] = A[r] * v[0] + A[3 + r] * v[1] + A[6 + r] * v[2];
}
__device__ __forceinline__ void col3(const double* A, int j, double* v) { v[0] = A[j]; v[1] = A[3 + j]; v[2] = A[6 + j]; }
__device__ __forceinline__ void cross3d(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3d(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// math::exp(Vector3), Utils.h:111-127
__device__ void exp3(const double* w, double* R) {
    const double wlen = sqrt(dot3d(w, w));
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (wlen > 1e-12) {
        const double wx = w[0] / wlen, wy = w[1] / wlen, wz = w[2] / wlen;
        const double c = cos(wlen), s = sin(wlen), c1 = 1.0 - c;
        R[0] = c + wx * wx * c1;       R[1] = -wz * s + wx * wy * c1; R[2] = wy * s + wx * wz * c1;
        R[3] = wz * s + wx * wy * c1;  R[4] = c + wy * wy * c1;       R[5] = -wx * s + wy * wz * c1;
        R[6] = -wy * s + wx * wz * c1; R[7] = wx * s + wy * wz * c1;  R[8] = c + wz * wz * c1;
    }
}

// JointSphericalExp::update (JointSphericalExp.cpp:13-183). T = _S_j top 3x3, Tdot = _S_j_dot top 3x3.
struct SphKin {
    double R[9], Rdot[9], T[9], Tdot[9];
    double dR[3][9], dRdot[3][9], dT[3][9], dTdot[3][9];
};

__device__ __noinline__ void spherical_exp(const double* q, const double* qd, bool deriv, SphKin& o) {
    double e[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double t9[9], u9[9], v3[3], w3[3];
    if (sqrt(dot3d(q, q)) < kMathEps) {  // case 1 (.cpp:27-66)
        for (int i = 0; i < 9; ++i) { o.R[i] = (i % 4 == 0) ? 1.0 : 0.0; o.Rdot[i] = 0; o.T[i] = o.R[i]; }
        for (int j = 0; j < 3; ++j) {
            skew(e[j], o.dR[j]);
            for (int i = 0; i < 9; ++i) { o.Rdot[i] += o.dR[j][i] * qd[j]; o.dRdot[j][i] = 0; o.dTdot[j][i] = 0; }
        }
        for (int j = 0; j < 3; ++j) {
            skew(e[j], t9);
            tmul33(o.Rdot, t9, u9);
            unskew(u9, v3);
            for (int r = 0; r < 3; ++r) o.Tdot[r * 3 + j] = v3[r];
            for (int k = 0; k < 3; ++k) {
                double sk[9];
                skew(e[k], sk);
                tmul33(sk, t9, u9);
                unskew(u9, v3);
                for (int r = 0; r < 3; ++r) o.dT[k][r * 3 + j] = v3[r];
            }
        }
        return;
    }
    exp3(q, o.R);
    const double q2 = dot3d(q, q), d = 1.0 / q2;
    double Sq[9], Sqd[9], ImR[9];
    skew(q, Sq); skew(qd, Sqd);
    for (int i = 0; i < 9; ++i) ImR[i] = ((i % 4 == 0) ? 1.0 : 0.0) - o.R[i];
    double A[3][9], BC[3][9];  // BC = B + C
    for (int j = 0; j < 3; ++j) {
        col3(ImR, j, v3);
        mulv3(Sq, v3, w3);
        skew(w3, t9);  // C[j]
        for (int i = 0; i < 9; ++i) { BC[j][i] = q[j] * Sq[i] + t9[i]; A[j][i] = BC[j][i] * d; }
        mul33(A[j], o.R, o.dR[j]);
    }
    for (int i = 0; i < 9; ++i) o.Rdot[i] = o.dR[0][i] * qd[0] + o.dR[1][i] * qd[1] + o.dR[2][i] * qd[2];
    const double qqd = dot3d(q, qd);
    const double ddot = -2.0 * qqd / (q2 * q2);
    double Adot[3][9], BCdot[3][9];
    for (int j = 0; j < 3; ++j) {
        double a3[3], b3[3];
        col3(ImR, j, v3);
        mulv3(Sqd, v3, a3);
        col3(o.Rdot, j, v3);
        mulv3(Sq, v3, b3);
        for (int r = 0; r < 3; ++r) w3[r] = a3[r] - b3[r];
        skew(w3, t9);  // Cdot[j]
        for (int i = 0; i < 9; ++i) {
            BCdot[j][i] = qd[j] * Sq[i] + q[j] * Sqd[i] + t9[i];
            Adot[j][i] = BCdot[j][i] * d + BC[j][i] * ddot;
        }
    }
    for (int j = 0; j < 3; ++j) {
        tmul33(o.R, o.dR[j], t9);
        unskew(t9, v3);
        for (int r = 0; r < 3; ++r) o.T[r * 3 + j] = v3[r];
        // Rdot^T A R + R^T Adot R + R^T A Rdot
        double m1[9], m2[9], acc[9];
        mul33(A[j], o.R, m1); tmul33(o.Rdot, m1, acc);
        mul33(Adot[j], o.R, m1); tmul33(o.R, m1, m2);
        for (int i = 0; i < 9; ++i) acc[i] += m2[i];
        mul33(A[j], o.Rdot, m1); tmul33(o.R, m1, m2);
        for (int i = 0; i < 9; ++i) acc[i] += m2[i];
        unskew(acc, v3);
        for (int r = 0; r < 3; ++r) o.Tdot[r * 3 + j] = v3[r];
    }
    if (!deriv) return;
    double dd_dq[3], dddot_dq[3];
    for (int k = 0; k < 3; ++k) {
        dd_dq[k] = -2.0 * q[k] / (q2 * q2);
        dddot_dq[k] = (-2.0 * qd[k] * q2 + 8.0 * qqd * q[k]) / (q2 * q2 * q2);
    }
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 9; ++i) o.dRdot[k][i] = 0;
    // dBC[j][k] = dB + dC, dA[j][k]; kept for the Tdot derivatives: 2 x 9 x 9 doubles
    double dA[3][3][9], dBC[3][3][9];
    for (int k = 0; k < 3; ++k) {
        double Sek[9];
        skew(e[k], Sek);
        for (int j = 0; j < 3; ++j) {
            double a3[3], b3[3];
            col3(ImR, j, v3);
            mulv3(Sek, v3, a3);      // skew(e_k) (I-R) e_j
            col3(o.dR[k], j, v3);
            mulv3(Sq, v3, b3);       // skew(q) dR_k e_j
            for (int r = 0; r < 3; ++r) w3[r] = a3[r] - b3[r];
            skew(w3, t9);            // dC[j][k]
            for (int i = 0; i < 9; ++i) {
                double db = q[j] * Sek[i];
                if (j == k) db += Sq[i];
                dBC[j][k][i] = db + t9[i];
                dA[j][k][i] = dBC[j][k][i] * d + BC[j][i] * dd_dq[k];
            }
            // d2R/dqj dqk = dA[j][k] R + A[j] dR[k]
            double m1[9], m2[9];
            mul33(dA[j][k], o.R, m1);
            mul33(A[j], o.dR[k], m2);
            for (int i = 0; i < 9; ++i) o.dRdot[k][i] += (m1[i] + m2[i]) * qd[j];
        }
    }
    for (int j = 0; j < 3; ++j) {
        for (int k = 0; k < 3; ++k) {
            double Sek[9], m1[9], m2[9], d2R[9], acc[9];
            skew(e[k], Sek);
            mul33(dA[j][k], o.R, m1);
            mul33(A[j], o.dR[k], m2);
            for (int i = 0; i < 9; ++i) d2R[i] = m1[i] + m2[i];
            // dT[k].col(j) = unskew(dR_k^T dR_j + R^T d2R)
            tmul33(o.dR[k], o.dR[j], m1);
            tmul33(o.R, d2R, m2);
            for (int i = 0; i < 9; ++i) acc[i] = m1[i] + m2[i];
            unskew(acc, v3);
            for (int r = 0; r < 3; ++r) o.dT[k][r * 3 + j] = v3[r];
            // dAdot[j][k]
            double a3[3], b3[3], c3[3], dAd[9];
            col3(o.dR[k], j, v3);  mulv3(Sqd, v3, a3);
            col3(o.Rdot, j, v3);   mulv3(Sek, v3, b3);
            col3(o.dRdot[k], j, v3); mulv3(Sq, v3, c3);
            for (int r = 0; r < 3; ++r) w3[r] = -a3[r] - b3[r] - c3[r];
            skew(w3, t9);  // dCdot
            for (int i = 0; i < 9; ++i) {
                double dbd = qd[j] * Sek[i];
                if (j == k) dbd += Sqd[i];
                dAd[i] = (dbd + t9[i]) * d + BCdot[j][i] * dd_dq[k] + dBC[j][k][i] * ddot + BC[j][i] * dddot_dq[k];
            }
            // dTdot[k].col(j): nine products (.cpp:166-169)
            for (int i = 0; i < 9; ++i) acc[i] = 0;
#define SDFC_ACC3(LT, M, RT)                    \
    do {                                        \
        mul33(M, RT, m1);                       \
        tmul33(LT, m1, m2);                     \
        for (int i = 0; i < 9; ++i) acc[i] += m2[i]; \
    } while (0)
            SDFC_ACC3(o.dRdot[k], A[j], o.R);
            SDFC_ACC3(o.Rdot, dA[j][k], o.R);
            SDFC_ACC3(o.Rdot, A[j], o.dR[k]);
            SDFC_ACC3(o.dR[k], Adot[j], o.R);
            SDFC_ACC3(o.R, dAd, o.R);
            SDFC_ACC3(o.R, Adot[j], o.dR[k]);
            SDFC_ACC3(o.dR[k], A[j], o.Rdot);
            SDFC_ACC3(o.R, dA[j][k], o.Rdot);
            SDFC_ACC3(o.R, A[j], o.dRdot[k]);
#undef SDFC_ACC3
            unskew(acc, v3);
            for (int r = 0; r < 3; ++r) o.dTdot[k][r * 3 + j] = v3[r];
        }
    }
}

// Joint quantities of one non-fixed body: S_j, S_j_dot (6 x n, row-major with row
// stride 6) and their q-derivatives (nonzero only for the rotational dofs 3..5 of
// a free3d-exp joint), plus Q = [R p; 0 1].
struct JointKin {
    double R[9], p[3];
    double S[36], Sd[36];
    double dS[3][36], dSd[3][36];  // derivative w.r.t. dof 3 + i
};

template <bool kFree>
__device__ __noinline__ void joint_kin(const BodyDev& b, const double* q, const double* qd, bool deriv, JointKin& k,
                                       SphKin& sph) {
    for (int i = 0; i < 36; ++i) { k.S[i] = 0; k.Sd[i] = 0; }
    for (int i = 0; i < 9; ++i) k.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    k.p[0] = k.p[1] = k.p[2] = 0;
    if (b.type == SDFC_JOINT_TRANSLATIONAL) {  // JointTranslational.cpp:9-40
        for (int a = 0; a < 3; ++a) { k.p[a] = q[a]; k.S[(3 + a) * 6 + a] = 1.0; }
    } else if (kFree && b.type == SDFC_JOINT_FREE3D_EXP) {  // JointFree3DExp.cpp:13-110
        spherical_exp(q + 3, qd + 3, deriv, sph);
        for (int i = 0; i < 9; ++i) k.R[i] = sph.R[i];
        for (int a = 0; a < 3; ++a) k.p[a] = q[a];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                k.S[r * 6 + 3 + c] = sph.T[r * 3 + c];
                k.S[(3 + r) * 6 + c] = sph.R[c * 3 + r];
                k.Sd[r * 6 + 3 + c] = sph.Tdot[r * 3 + c];
                k.Sd[(3 + r) * 6 + c] = sph.Rdot[c * 3 + r];
            }
        if (deriv)
            for (int i = 0; i < 3; ++i) {
                for (int e = 0; e < 36; ++e) { k.dS[i][e] = 0; k.dSd[i][e] = 0; }
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        k.dS[i][r * 6 + 3 + c] = sph.dT[i][r * 3 + c];
                        k.dS[i][(3 + r) * 6 + c] = sph.dR[i][c * 3 + r];
                        k.dSd[i][r * 6 + 3 + c] = sph.dTdot[i][r * 3 + c];
                        k.dSd[i][(3 + r) * 6 + c] = sph.dRdot[i][c * 3 + r];
                    }
            }
    }
}

// E_0i = E_pj_0 * Q * E_ji (Joint.cpp:101-113, Body.cpp:140), written column-major;
// phi_i = A_ij (S_j qdot) (Joint.cpp:119-121, Body.cpp:156).
__device__ void body_pose(const BodyDev& b, const JointKin& k, const double* qd, double* E16, double* phi6, double* R0i) {
    double R0j[9], p0j[3], t3[3];
    mul33(b.Rp, k.R, R0j);
    mulv3(b.Rp, k.p, t3);
    for (int a = 0; a < 3; ++a) p0j[a] = t3[a] + b.pp[a];
    double R[9], p[3];
    mul33(R0j, b.Rji, R);
    mulv3(R0j, b.pji, t3);
    for (int a = 0; a < 3; ++a) p[a] = t3[a] + p0j[a];
    if (E16) {
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) E16[r + 4 * c] = R[r * 3 + c];
            E16[r + 12] = p[r];
            E16[3 + 4 * r] = 0.0;
        }
        E16[15] = 1.0;
    }
    if (R0i) for (int i = 0; i < 9; ++i) R0i[i] = R[i];
    if (phi6) {
        double pj[6];
        for (int r = 0; r < 6; ++r) {
            double s = 0;
            for (int c = 0; c < b.ndof; ++c) s += k.S[r * 6 + c] * qd[c];
            pj[r] = s;
        }
        for (int r = 0; r < 6; ++r) {
            double s = 0;
            for (int c = 0; c < 6; ++c) s += b.Aij[r * 6 + c] * pj[c];
            phi6[r] = (b.ndof > 0) ? s : 0.0;
        }
    }
}

// ------------------------------------------------------------------ kernels
__global__ void k_sim_begin(SimDev S, int n_steps) {
    const int env = S.env0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= S.env0 + S.env_count) return;
    // environments masked out by sdfc_sim_set_active keep their state untouched
    const bool on = n_steps > 0 && (S.active == nullptr || S.active[env] != 0);
    if (on) atomicAdd(S.n_unfinished, 1);
    S.steps_left[env] = on ? n_steps : 0;
    S.step_idx[env] = 0;
    S.phase[env] = on ? PH_DERIV : PH_DONE;
    S.iter[env] = 0;
    S.strike[env] = 0;
    const size_t o = (size_t)env * S.nr;
    for (int j = 0; j < S.nr; ++j) {  // integration_BDF1: q1 = q0 + h qdot0 (Simulation.cpp:1255)
        S.q0[o + j] = S.q[o + j];
        S.qdot0[o + j] = S.qdot[o + j];
        S.x[o + j] = S.q[o + j] + S.h * S.qdot[o + j];
    }
}

// candidate state of this round: x (Newton evaluation) or x + alpha dx (line-search trial)
__device__ __forceinline__ void candidate(const SimDev& S, int env, int phase, double* xc, double* qdc) {
    const size_t o = (size_t)env * S.nr;
    const double alpha = S.alpha[env];
    for (int j = 0; j < S.nr; ++j) {
        const double x = (phase == PH_LS) ? S.x[o + j] + alpha * S.dx[o + j] : S.x[o + j];
        xc[j] = x;
        qdc[j] = (x - S.q0[o + j]) / S.h;  // Simulation.cpp:1082,1095
    }
}

template <int kMaxR, int kMaxDyn, bool kFree>
__global__ void __launch_bounds__(32) k_sim_pre(SimDev S, int poses_only) {
    const int env = S.env0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= S.env0 + S.env_count) return;
    const int phase = poses_only ? PH_DERIV : S.phase[env];
    if (!poses_only) S.env_mode[env] = (phase == PH_DONE) ? 0 : (phase == PH_LS ? 1 : 2);
    if (phase == PH_DONE) return;
    double xc[kMaxR], qdc[kMaxR];
    if (poses_only) {
        for (int j = 0; j < S.nr; ++j) { xc[j] = S.q[(size_t)env * S.nr + j]; qdc[j] = S.qdot[(size_t)env * S.nr + j]; }
    } else {
        candidate(S, env, phase, xc, qdc);
    }
    JointKin jk;
    SphKin sph;
    for (int b = 0; b < S.n_bodies; ++b) {
        const BodyDev& B = S.bodies[b];
        joint_kin<kFree>(B, xc + B.dof_off, qdc + B.dof_off, false, jk, sph);
        body_pose(B, jk, qdc + B.dof_off, S.E + ((size_t)env * S.n_bodies + b) * 16,
                  S.phi + ((size_t)env * S.n_bodies + b) * 6, nullptr);
    }
}

// Solve A x = b (n <= kMaxR) by LU with partial pivoting (Simulation.cpp:1028 uses Eigen's PartialPivLU).
template <int kMaxR>
__device__ __noinline__ void lu_solve(double* A, double* b, int n) {
    for (int c = 0; c < n; ++c) {
        int piv = c;
        double best = fabs(A[c * kMaxR + c]);
        for (int r = c + 1; r < n; ++r) {
            const double v = fabs(A[r * kMaxR + c]);
            if (v > best) { best = v; piv = r; }
        }
        if (piv != c) {
            for (int k = 0; k < n; ++k) { const double t = A[c * kMaxR + k]; A[c * kMaxR + k] = A[piv * kMaxR + k]; A[piv * kMaxR + k] = t; }
            const double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        const double inv = 1.0 / A[c * kMaxR + c];
        for (int r = c + 1; r < n; ++r) {
            const double f = A[r * kMaxR + c] * inv;
            A[r * kMaxR + c] = f;
            for (int k = c + 1; k < n; ++k) A[r * kMaxR + k] -= f * A[c * kMaxR + k];
            b[r] -= f * b[c];
        }
    }
    for (int r = n - 1; r >= 0; --r) {
        double s = b[r];
        for (int k = r + 1; k < n; ++k) s -= A[r * kMaxR + k] * b[k];
        b[r] = s / A[r * kMaxR + r];
    }
}

// Per non-fixed body: Jacobian blocks and maximal-coordinate force data.
struct DynBody {
    double J[36], Jd[36];            // 6 x n (row stride 6)
    double dJ[3][36], dJd[3][36];    // w.r.t. rotational dofs (free3d-exp only)
    double fm[6];
    int n, off, body;
    bool has_dJ;
};

__device__ __forceinline__ void mul66n(const double* A, const double* B, int n, double* C) {  // C(6xn) = A(6x6) B(6xn)
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < n; ++c) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += A[r * 6 + k] * B[k * 6 + c];
            C[r * 6 + c] = s;
        }
}

template <int kMaxR, int kMaxDyn, bool kFree>
__global__ void __launch_bounds__(32) k_sim_post(SimDev S) {
    const int env = S.env0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= S.env0 + S.env_count) return;
    const int mode = S.env_mode[env];
    if (mode == 0) return;
    const bool deriv = (mode == 2);
    const int nr = S.nr;
    const double h = S.h;
    const size_t o = (size_t)env * nr;
    double xc[kMaxR], qdc[kMaxR];
    candidate(S, env, deriv ? PH_DERIV : PH_LS, xc, qdc);
    atomicAdd(S.counters + (deriv ? 0 : 1), 1ull);

    // contact history (CollisionDetection.cpp:143-146)
    for (int p = 0; p < S.n_pairs; ++p)
        if (S.n_active[(size_t)env * S.n_pairs + p] > 0) S.touched[(size_t)env * S.n_pairs + p] = 1;

    DynBody dyn[kMaxDyn];
    double Kmm[kMaxDyn][kMaxDyn][36], Dmm[kMaxDyn][kMaxDyn][36];  // maximal blocks between non-fixed bodies
    if (deriv)
        for (int s = 0; s < kMaxDyn; ++s)
            for (int t = 0; t < kMaxDyn; ++t)
                for (int i = 0; i < 36; ++i) { Kmm[s][t][i] = 0; Dmm[s][t][i] = 0; }

    // ---- kinematics + body forces of the non-fixed bodies (Body.cpp:250-330)
    {
        JointKin jk;
        SphKin sph;
        for (int b = 0; b < S.n_bodies; ++b) {
            const BodyDev& B = S.bodies[b];
            if (B.dyn < 0) continue;
            DynBody& D = dyn[B.dyn];
            D.n = B.ndof; D.off = B.dof_off; D.body = b;
            D.has_dJ = kFree && deriv && (B.type == SDFC_JOINT_FREE3D_EXP);
            joint_kin<kFree>(B, xc + B.dof_off, qdc + B.dof_off, deriv, jk, sph);
            double R[9], phi[6];
            body_pose(B, jk, qdc + B.dof_off, nullptr, phi, R);
            mul66n(B.Aij, jk.S, D.n, D.J);     // Robot.cpp:664-665
            mul66n(B.Aij, jk.Sd, D.n, D.Jd);
            if (kFree && D.has_dJ)
                for (int i = 0; i < 3; ++i) { mul66n(B.Aij, jk.dS[i], D.n, D.dJ[i]); mul66n(B.Aij, jk.dSd[i], D.n, D.dJd[i]); }
            // Coriolis: ad(phi)^T (I phi)
            const double* w = phi; const double* v = phi + 3;
            double Iw[3], mv[3], c1[3], c2[3], c3[3];
            for (int a = 0; a < 3; ++a) { Iw[a] = B.I6[a] * w[a]; mv[a] = B.I6[3 + a] * v[a]; }
            cross3d(w, Iw, c1); cross3d(v, mv, c2); cross3d(w, mv, c3);
            // ad^T = [[w]^T [v]^T; 0 [w]^T] -> top: -(w x Iw) - (v x mv), bottom: -(w x mv)
            const double* fe = S.fext + ((size_t)env * S.n_bodies + b) * 6;
            double ge[3] = {B.mass * S.grav[0], B.mass * S.grav[1], B.mass * S.grav[2]}, fg[3], e1[3], e2[3];
            tmulv3(R, ge, fg);
            tmulv3(R, fe, e1);
            tmulv3(R, fe + 3, e2);
            for (int a = 0; a < 3; ++a) {
                D.fm[a] = (-c1[a] - c2[a]) + B.I6[a] * e1[a];
                D.fm[3 + a] = -c3[a] + fg[a] + B.I6[3 + a] * e2[a];
            }
            if (deriv) {
                double* K = Kmm[B.dyn][B.dyn];
                double* Dm = Dmm[B.dyn][B.dyn];
                double Sg[9], S1[9], S2[9];
                skew(fg, Sg); skew(e1, S1); skew(e2, S2);
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        K[(3 + r) * 6 + c] += Sg[r * 3 + c] + S2[r * 3 + c] * B.I6[3 + r];
                        K[r * 6 + c] += S1[r * 3 + c] * B.I6[r];
                    }
                // Dm = ad^T diag(I) - [ [Iw] [mv]; [mv] 0 ]   (Body.cpp:305-329)
                double Sw[9], Sv[9], SIw[9], Smv[9];
                skew(w, Sw); skew(v, Sv); skew(Iw, SIw); skew(mv, Smv);
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        // ad^T blocks: TL = [w]^T, TR = [v]^T, BR = [w]^T; column k scaled by I(k)
                        Dm[r * 6 + c] += Sw[c * 3 + r] * B.I6[c];
                        Dm[r * 6 + 3 + c] += Sv[c * 3 + r] * B.I6[3 + c];
                        Dm[(3 + r) * 6 + 3 + c] += Sw[c * 3 + r] * B.I6[3 + c];
                        // minus e_k x Iw etc.: column k of -[Iw]^T... (skew(e_k) a = -skew(a) e_k)
                        Dm[r * 6 + c] += SIw[r * 3 + c];
                        Dm[r * 6 + 3 + c] += Smv[r * 3 + c];
                        Dm[(3 + r) * 6 + c] += Smv[r * 3 + c];
                    }
            }
        }
    }
    // ---- contact blocks, in force order (Robot.cpp:524-525, 546-547)
    for (int p = 0; p < S.n_pairs; ++p) {
        const size_t ep = (size_t)env * S.n_pairs + p;
        const int sl[2] = {S.bodies[S.pair_a[p]].dyn, S.bodies[S.pair_b[p]].dyn};
        for (int br = 0; br < 2; ++br) {
            if (sl[br] < 0) continue;
            for (int r = 0; r < 6; ++r) dyn[sl[br]].fm[r] += S.fm_sub[ep * 12 + 6 * br + r];
            if (!deriv) continue;
            for (int bc = 0; bc < 2; ++bc) {
                if (sl[bc] < 0) continue;
                double* K = Kmm[sl[br]][sl[bc]];
                double* Dm = Dmm[sl[br]][sl[bc]];
                for (int r = 0; r < 6; ++r)
                    for (int c = 0; c < 6; ++c) {
                        const size_t src = ep * 144 + (6 * br + r) + 12 * (6 * bc + c);  // column-major 12x12
                        K[r * 6 + c] += S.Km_sub[src];
                        Dm[r * 6 + c] += S.Dm_sub[src];
                    }
            }
        }
    }

    // ---- ground contact blocks (ForceGroundContact.cpp:129,186-189): body-diagonal
    for (int gi = 0; gi < S.n_grounds; ++gi) {
        const size_t eg = (size_t)env * S.n_grounds + gi;
        if (S.g_n_active[eg] > 0) S.touched_ground[eg] = 1;  // CollisionDetection.cpp:67
        const int sl = S.bodies[S.ground_body[gi]].dyn;
        if (sl < 0) continue;
        const double* o = S.g_out + eg * 78;
        for (int r = 0; r < 6; ++r) dyn[sl].fm[r] += o[r];
        if (!deriv) continue;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                Kmm[sl][sl][r * 6 + c] += o[6 + r + 6 * c];
                Dmm[sl][sl][r * 6 + c] += o[42 + r + 6 * c];
            }
    }

    // ---- reduced assembly (Simulation.cpp:379-503) and BDF1 residual (:1081-1112)
    double g[kMaxR], H[kMaxR * kMaxR];
    for (int i = 0; i < nr; ++i) {
        g[i] = 0;
        if (deriv) for (int j = 0; j < nr; ++j) H[i * kMaxR + j] = 0;
    }
    double dJq[kMaxDyn][36];  // dJ_dq(k) * qdot, 6 x n per body
    for (int s = 0; s < S.n_dyn; ++s) {
        DynBody& D = dyn[s];
        const BodyDev& B = S.bodies[D.body];
        const int n = D.n, off = D.off;
        const double* qd = qdc + off;
        double MmJ[36], M[36], JTMmJd[36], dq[6], MmJdq[6];
        for (int r = 0; r < 6; ++r) {
            double s2 = 0;
            for (int c = 0; c < n; ++c) { MmJ[r * 6 + c] = D.J[r * 6 + c] * B.I6[r]; s2 += D.Jd[r * 6 + c] * B.I6[r] * qd[c]; }
            MmJdq[r] = s2;
        }
        for (int i = 0; i < n; ++i) {
            dq[i] = xc[off + i] - S.q0[o + off + i] - h * S.qdot0[o + off + i];
            for (int j = 0; j < n; ++j) {
                double m = 0, jj = 0;
                for (int r = 0; r < 6; ++r) { m += D.J[r * 6 + i] * MmJ[r * 6 + j]; jj += MmJ[r * 6 + i] * D.Jd[r * 6 + j]; }
                M[i * 6 + j] = m;
                JTMmJd[i * 6 + j] = jj;
            }
        }
        // f = J^T fm - J^T Mm Jdot qdot ; g = M dq - h^2 f
        for (int i = 0; i < n; ++i) {
            double f = 0, md = 0;
            for (int r = 0; r < 6; ++r) f += D.J[r * 6 + i] * (D.fm[r] - MmJdq[r]);
            for (int j = 0; j < n; ++j) md += M[i * 6 + j] * dq[j];
            g[off + i] = md - h * h * f;
        }
        if (!deriv) continue;
        // dJq(:,k) = dJ_dq(k) qdot
        for (int r = 0; r < 6; ++r)
            for (int k = 0; k < n; ++k) {
                double s2 = 0;
                if (kFree && D.has_dJ && k >= 3)
                    for (int c = 0; c < n; ++c) s2 += D.dJ[k - 3][r * 6 + c] * qd[c];
                dJq[s][r * 6 + k] = s2;
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                // H = M - h^2 K - h D : the M, Kqvv (below) and Dqvv parts of the diagonal block
                double t = 0;
                for (int r = 0; r < 6; ++r) t += MmJ[r * 6 + i] * dJq[s][r * 6 + j];
                const double Dqvv = -JTMmJd[i * 6 + j] - t;
                H[(off + i) * kMaxR + off + j] += M[i * 6 + j] - h * Dqvv;
            }
        if (kFree && D.has_dJ) {
            for (int k = 3; k < n; ++k) {
                const double* dJk = D.dJ[k - 3];
                const double* dJdk = D.dJd[k - 3];
                double v6[6];  // dJdot_dq(k) qdot
                for (int r = 0; r < 6; ++r) {
                    double s2 = 0;
                    for (int c = 0; c < n; ++c) s2 += dJdk[r * 6 + c] * qd[c];
                    v6[r] = s2;
                }
                for (int i = 0; i < n; ++i) {
                    // Kqvv(:,k) = -dJ_k^T MmJdq - JTMm v6 ; plus dJ_k^T fm ; dM_dq(k) dq
                    double a = 0, b2 = 0, c2 = 0, dm = 0;
                    for (int r = 0; r < 6; ++r) {
                        a += dJk[r * 6 + i] * MmJdq[r];
                        b2 += MmJ[r * 6 + i] * v6[r];
                        c2 += dJk[r * 6 + i] * D.fm[r];
                    }
                    for (int j = 0; j < n; ++j) {
                        double t1 = 0, t2 = 0;  // tmp(i,j) = dJk^T MmJ ; tmp(j,i)
                        for (int r = 0; r < 6; ++r) { t1 += dJk[r * 6 + i] * MmJ[r * 6 + j]; t2 += dJk[r * 6 + j] * MmJ[r * 6 + i]; }
                        dm += (t1 + t2) * dq[j];
                    }
                    H[(off + i) * kMaxR + off + k] += -h * h * ((-a - b2) + c2) + dm;
                }
            }
        }
    }
    if (deriv) {
        // cross terms: J_s^T Km_st J_t, J_s^T Dm_st dJq_t (into K) and J_s^T Dm_st J_t (into D)
        for (int s = 0; s < S.n_dyn; ++s)
            for (int t = 0; t < S.n_dyn; ++t) {
                const DynBody& A = dyn[s];
                const DynBody& Bt = dyn[t];
                double JTK[36], JTD[36];  // n_s x 6
                for (int i = 0; i < A.n; ++i)
                    for (int c = 0; c < 6; ++c) {
                        double a = 0, b2 = 0;
                        for (int r = 0; r < 6; ++r) { a += A.J[r * 6 + i] * Kmm[s][t][r * 6 + c]; b2 += A.J[r * 6 + i] * Dmm[s][t][r * 6 + c]; }
                        JTK[i * 6 + c] = a; JTD[i * 6 + c] = b2;
                    }
                for (int i = 0; i < A.n; ++i)
                    for (int j = 0; j < Bt.n; ++j) {
                        double k1 = 0, k2 = 0, d1 = 0;
                        for (int c = 0; c < 6; ++c) {
                            k1 += JTK[i * 6 + c] * Bt.J[c * 6 + j];
                            k2 += JTD[i * 6 + c] * dJq[t][c * 6 + j];
                            d1 += JTD[i * 6 + c] * Bt.J[c * 6 + j];
                        }
                        H[(A.off + i) * kMaxR + Bt.off + j] += -h * h * (k1 + k2) - h * d1;
                    }
            }
    }

    double gn = 0;
    for (int i = 0; i < nr; ++i) gn += g[i] * g[i];
    gn = sqrt(gn);

    // ---- Newton / line-search state machine (Simulation.cpp:999-1079)
    if (deriv) {
        double rhs[kMaxR];
        for (int i = 0; i < nr; ++i) rhs[i] = -g[i];
        lu_solve<kMaxR>(H, rhs, nr);
        for (int i = 0; i < nr; ++i) S.dx[o + i] = rhs[i];
        S.gnorm[env] = gn;
        S.alpha[env] = 1.0;
        S.trial[env] = 0;
        S.iter[env] += 1;
        S.phase[env] = PH_LS;
        return;
    }
    bool accept = false, end_newton = false, success = false;
    if (gn < S.gnorm[env]) {  // success_ls
        S.strike[env] = 0;
        accept = true;
    } else {
        const int trial = S.trial[env] + 1;
        S.trial[env] = trial;
        S.alpha[env] *= 0.5;
        if (trial < kMaxLS) return;  // next trial: phase stays PH_LS
        const int strike = S.strike[env] + 1;
        S.strike[env] = strike;
        if (strike >= kMaxStrike) end_newton = true;  // break before applying the step
        else accept = true;                            // applied with alpha = 0.5^20
    }
    if (accept) {
        const double alpha = S.alpha[env];
        for (int i = 0; i < nr; ++i) S.x[o + i] += alpha * S.dx[o + i];
        if (gn < kNewtonTol) { success = true; end_newton = true; }
        else if (S.iter[env] >= S.max_iter) end_newton = true;
        else S.phase[env] = PH_DERIV;
    }
    if (!end_newton) return;

    // ---- end of the step (Simulation.cpp:955-966, 1259): qdot1, reparam, history
    S.converged[env] = success ? 1 : 0;
    double q1[kMaxR], qd1[kMaxR];
    for (int i = 0; i < nr; ++i) { q1[i] = S.x[o + i]; qd1[i] = (q1[i] - S.q0[o + i]) / h; }
    for (int b = 0; b < S.n_bodies; ++b) {
        const BodyDev& B = S.bodies[b];
        if (B.type != SDFC_JOINT_FREE3D_EXP) continue;
        double* r = q1 + B.dof_off + 3;
        double* rd = qd1 + B.dof_off + 3;
        double qn = sqrt(dot3d(r, r));
        const double pi = acos(-1.0);
        if (qn > 1.5 * pi) {  // JointSphericalExp::reparam (.cpp:237-264)
            SphKin sph;
            spherical_exp(r, rd, false, sph);
            double S0rd[3];
            mulv3(sph.T, rd, S0rd);
            while (qn > 1.5 * pi) {
                const double f = 1.0 - 2.0 * pi / qn;
                for (int a = 0; a < 3; ++a) r[a] *= f;
                qn = sqrt(dot3d(r, r));
            }
            spherical_exp(r, rd, false, sph);
            double A3[kMaxR * kMaxR], b3[kMaxR];
            for (int i = 0; i < 3; ++i) { b3[i] = S0rd[i]; for (int j = 0; j < 3; ++j) A3[i * kMaxR + j] = sph.T[i * 3 + j]; }
            lu_solve<kMaxR>(A3, b3, 3);
            for (int a = 0; a < 3; ++a) rd[a] = b3[a];
        }
    }
    const int step = S.step_idx[env];
    for (int i = 0; i < nr; ++i) {
        S.q[o + i] = q1[i];
        S.qdot[o + i] = qd1[i];
        if (S.q_traj) S.q_traj[((size_t)step * S.n_envs + env) * nr + i] = q1[i];
        if (S.qdot_traj) S.qdot_traj[((size_t)step * S.n_envs + env) * nr + i] = qd1[i];
    }
    atomicAdd(S.counters + 2, 1ull);
    S.step_idx[env] = step + 1;
    const int left = S.steps_left[env] - 1;
    S.steps_left[env] = left;
    if (left <= 0) {
        S.phase[env] = PH_DONE;
        atomicSub(S.n_unfinished, 1);
        return;
    }
    for (int i = 0; i < nr; ++i) {
        S.q0[o + i] = q1[i];
        S.qdot0[o + i] = qd1[i];
        S.x[o + i] = q1[i] + h * qd1[i];
    }
    S.iter[env] = 0;
    S.strike[env] = 0;
    S.phase[env] = PH_DERIV;
}

}  // namespace sdfc

// ============================================================================
// host side of the sdfc_sim_* entry points
// ============================================================================
using namespace sdfc;
using sdfc_host::dev_alloc;
using sdfc_host::fail;

struct sdfc_sim {
    sdfc_batch* batch = nullptr;
    sdfc_context* ctx = nullptr;
    SimDev dev{};
    int n_envs = 0, n_bodies = 0, n_pairs = 0, nr = 0;
    BodyDev* d_bodies = nullptr;
    int *d_pair_a = nullptr, *d_pair_b = nullptr;
    double* d_state = nullptr;   // q, qdot, q0, qdot0, x, dx
    double* d_scalars = nullptr; // gnorm, alpha
    int* d_ints = nullptr;       // phase, trial, iter, strike, steps_left, step_idx, env_mode
    uint8_t *d_converged = nullptr, *d_touched = nullptr, *d_active = nullptr, *d_touched_ground = nullptr;
    int* d_ground_body = nullptr;
    int n_grounds = 0;
    // lanes: disjoint environment ranges stepped on their own streams (see BatchDev::env0)
    static constexpr int kMaxLanes = 4;
    int n_lanes = 1;
    cudaStream_t lane_stream[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t lane_event[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork_event = nullptr;
    int variant = 2;  // kernel instantiation: 0 = (3,1,no free3d), 1 = (6,1,free3d), 2 = (12,2,free3d)
    bool use_active = false;
    double* d_fext = nullptr;
    int* d_unfinished = nullptr;
    unsigned long long* d_counters = nullptr;
    double *d_traj_q = nullptr, *d_traj_qd = nullptr;
    size_t traj_cap = 0;  // steps
    int* h_unfinished = nullptr;  // pinned
    int64_t rounds = 0;
};

namespace {
void colmajor_to_Rp(const double* E, double* R, double* p) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[r * 3 + c] = E[r + 4 * c];
        p[r] = E[r + 12];
    }
}
}  // namespace

namespace {
void launch_pre(sdfc_sim* s, const SimDev& D, int poses_only, int blocks, cudaStream_t st) {
    if (s->variant == 0) k_sim_pre<3, 1, false><<<blocks, 32, 0, st>>>(D, poses_only);
    else if (s->variant == 1) k_sim_pre<6, 1, true><<<blocks, 32, 0, st>>>(D, poses_only);
    else k_sim_pre<kCapR, kCapDyn, true><<<blocks, 32, 0, st>>>(D, poses_only);
}
void launch_post(sdfc_sim* s, const SimDev& D, int blocks, cudaStream_t st) {
    if (s->variant == 0) k_sim_post<3, 1, false><<<blocks, 32, 0, st>>>(D);
    else if (s->variant == 1) k_sim_post<6, 1, true><<<blocks, 32, 0, st>>>(D);
    else k_sim_post<kCapR, kCapDyn, true><<<blocks, 32, 0, st>>>(D);
}
}  // namespace

extern "C" {

sdfc_status sdfc_sim_create(sdfc_batch* batch, const sdfc_body_desc* bodies, const double gravity[3], double h,
                            sdfc_sim** out) {
    SDFC_REQUIRE(batch && bodies && gravity && out, "sdfc_sim_create: NULL argument");
    SDFC_REQUIRE(h > 0, "sdfc_sim_create: timestep must be positive");
    *out = nullptr;
    const int nb = batch->n_bodies, ne = batch->n_envs, np = batch->n_pairs;
    std::vector<BodyDev> hb(nb);
    int nr = 0, nd = 0;
    for (int b = 0; b < nb; ++b) {
        BodyDev& B = hb[b];
        const sdfc_body_desc& d = bodies[b];
        SDFC_REQUIRE(d.joint_type == SDFC_JOINT_FIXED || d.joint_type == SDFC_JOINT_TRANSLATIONAL ||
                     d.joint_type == SDFC_JOINT_FREE3D_EXP, "sdfc_sim_create: unsupported joint type");
        B.type = d.joint_type;
        B.ndof = d.joint_type == SDFC_JOINT_FIXED ? 0 : (d.joint_type == SDFC_JOINT_TRANSLATIONAL ? 3 : 6);
        B.dof_off = nr;
        B.dyn = B.ndof > 0 ? nd++ : -1;
        nr += B.ndof;
        colmajor_to_Rp(d.E_pj_0, B.Rp, B.pp);
        colmajor_to_Rp(d.E_ji, B.Rji, B.pji);
        // A_ij = Ad(E_ij), E_ij = Einv(E_ji): R = Rji^T, p = -Rji^T pji  (Utils.h:76-97)
        double R[9], p[3], Sp[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = B.Rji[c * 3 + r];
        for (int r = 0; r < 3; ++r) p[r] = -(R[r * 3] * B.pji[0] + R[r * 3 + 1] * B.pji[1] + R[r * 3 + 2] * B.pji[2]);
        Sp[0] = 0; Sp[1] = -p[2]; Sp[2] = p[1]; Sp[3] = p[2]; Sp[4] = 0; Sp[5] = -p[0]; Sp[6] = -p[1]; Sp[7] = p[0]; Sp[8] = 0;
        for (int i = 0; i < 36; ++i) B.Aij[i] = 0;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                B.Aij[r * 6 + c] = R[r * 3 + c];
                B.Aij[(3 + r) * 6 + 3 + c] = R[r * 3 + c];
                B.Aij[(3 + r) * 6 + c] = Sp[r * 3] * R[c] + Sp[r * 3 + 1] * R[3 + c] + Sp[r * 3 + 2] * R[6 + c];
            }
        for (int i = 0; i < 6; ++i) B.I6[i] = d.inertia[i];
        B.mass = d.mass;
    }
    SDFC_REQUIRE(nr <= kCapR && nd <= kCapDyn, "sdfc_sim_create: at most 2 non-fixed bodies / 12 reduced dofs are supported");
    SDFC_REQUIRE(nr > 0, "sdfc_sim_create: no degrees of freedom");
    SDFC_CUDA(cudaSetDevice(batch->ctx->device));
    sdfc_sim* s = new sdfc_sim();
    s->batch = batch; s->ctx = batch->ctx; s->n_envs = ne; s->n_bodies = nb; s->n_pairs = np; s->nr = nr;
    {
        bool any_free = false;
        for (int b = 0; b < nb; ++b) any_free |= hb[b].type == SDFC_JOINT_FREE3D_EXP;
        s->variant = (nd == 1 && nr <= 3 && !any_free) ? 0 : ((nd == 1 && nr <= 6) ? 1 : 2);
    }
    sdfc_status st = SDFC_OK;
    auto A = [&](sdfc_status x) { if (st == SDFC_OK) st = x; };
    A(dev_alloc(&s->d_bodies, (size_t)nb));
    A(dev_alloc(&s->d_pair_a, (size_t)np));
    A(dev_alloc(&s->d_pair_b, (size_t)np));
    A(dev_alloc(&s->d_state, (size_t)6 * ne * nr));
    A(dev_alloc(&s->d_scalars, (size_t)2 * ne));
    A(dev_alloc(&s->d_ints, (size_t)7 * ne));
    A(dev_alloc(&s->d_converged, (size_t)ne));
    A(dev_alloc(&s->d_touched, (size_t)ne * np));
    A(dev_alloc(&s->d_active, (size_t)ne));
    const int ng = batch->gset ? sdfc_host::ground_set_count(batch->gset) : 0;
    s->n_grounds = ng;
    A(dev_alloc(&s->d_touched_ground, (size_t)ne * ng));
    A(dev_alloc(&s->d_ground_body, (size_t)ng));
    A(dev_alloc(&s->d_fext, (size_t)ne * nb * 6));
    A(dev_alloc(&s->d_unfinished, (size_t)sdfc_sim::kMaxLanes));
    A(dev_alloc(&s->d_counters, (size_t)4));
    {   // SDFC_SKIN=0/1: temporal-coherence culling of the detection kernel (BatchDev::skin_*); results are identical
        const char* skin_env = std::getenv("SDFC_SKIN");
        if (skin_env ? std::atoi(skin_env) != 0 : (SDFC_SKIN_DEFAULT != 0)) A(sdfc_host::batch_enable_skin(batch));
    }
    if (st == SDFC_OK) {
        // SDFC_SIM_LANES overrides the lane count (1 = everything on the context's stream)
        const char* lanes_env = std::getenv("SDFC_SIM_LANES");
        s->n_lanes = lanes_env ? std::max(1, std::min((int)sdfc_sim::kMaxLanes, std::atoi(lanes_env))) : (ne >= 512 ? 2 : 1);
        s->n_lanes = std::min(s->n_lanes, std::max(1, ne / 32));
        cudaError_t e = cudaMallocHost((void**)&s->h_unfinished, sizeof(int) * sdfc_sim::kMaxLanes);
        for (int l = 0; l < s->n_lanes && e == cudaSuccess; ++l) {
            e = cudaStreamCreateWithFlags(&s->lane_stream[l], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->lane_event[l], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->fork_event, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_bodies, hb.data(), sizeof(BodyDev) * nb, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_pair_a, batch->body_a.data(), sizeof(int) * np, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_pair_b, batch->body_b.data(), sizeof(int) * np, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(s->d_state, 0, sizeof(double) * 6 * ne * nr);
        if (e == cudaSuccess) e = cudaMemset(s->d_scalars, 0, sizeof(double) * 2 * ne);
        if (e == cudaSuccess) e = cudaMemset(s->d_ints, 0, sizeof(int) * 7 * ne);
        if (e == cudaSuccess) e = cudaMemset(s->d_converged, 1, ne);
        if (e == cudaSuccess) e = cudaMemset(s->d_touched, 0, (size_t)ne * np);
        if (e == cudaSuccess && ng) e = cudaMemset(s->d_touched_ground, 0, (size_t)ne * ng);
        if (e == cudaSuccess && ng) e = cudaMemcpy(s->d_ground_body, sdfc_host::ground_set_bodies(batch->gset), sizeof(int) * ng, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(s->d_fext, 0, sizeof(double) * ne * nb * 6);
        if (e == cudaSuccess) e = cudaMemset(s->d_counters, 0, sizeof(unsigned long long) * 4);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();  // null-stream work done before the lane / context streams use it
        if (e != cudaSuccess) st = fail(SDFC_ERR_CUDA, std::string("sdfc_sim_create: ") + cudaGetErrorString(e));
    }
    if (st != SDFC_OK) { sdfc_sim_destroy(s); return st; }
    SimDev& D = s->dev;
    D.bodies = s->d_bodies; D.pair_a = s->d_pair_a; D.pair_b = s->d_pair_b;
    D.n_bodies = nb; D.n_pairs = np; D.n_envs = ne; D.nr = nr; D.n_dyn = nd;
    D.max_iter = std::max(20 * nr, 100);  // Simulation.cpp:1009
    D.h = h;
    for (int a = 0; a < 3; ++a) D.grav[a] = gravity[a];
    const size_t n = (size_t)ne * nr;
    D.q = s->d_state; D.qdot = s->d_state + n; D.q0 = s->d_state + 2 * n; D.qdot0 = s->d_state + 3 * n;
    D.x = s->d_state + 4 * n; D.dx = s->d_state + 5 * n;
    D.gnorm = s->d_scalars; D.alpha = s->d_scalars + ne;
    D.phase = s->d_ints; D.trial = s->d_ints + ne; D.iter = s->d_ints + 2 * ne; D.strike = s->d_ints + 3 * ne;
    D.steps_left = s->d_ints + 4 * ne; D.step_idx = s->d_ints + 5 * ne; D.env_mode = s->d_ints + 6 * ne;
    D.converged = s->d_converged; D.touched = s->d_touched; D.fext = s->d_fext; D.active = nullptr;
    D.E = batch->d_E; D.phi = batch->d_phi;
    D.fm_sub = batch->d_fm_sub; D.Km_sub = batch->d_Km_sub; D.Dm_sub = batch->d_Dm_sub; D.n_active = batch->d_n_active;
    D.n_grounds = ng; D.ground_body = s->d_ground_body; D.touched_ground = s->d_touched_ground;
    D.g_out = ng ? sdfc_host::ground_set_out(batch->gset) : nullptr;
    D.g_n_active = ng ? sdfc_host::ground_set_n_active(batch->gset) : nullptr;
    D.q_traj = nullptr; D.qdot_traj = nullptr;
    D.n_unfinished = s->d_unfinished; D.counters = s->d_counters;
    D.env0 = 0; D.env_count = ne;
    *out = s;
    return SDFC_OK;
}

sdfc_status sdfc_sim_destroy(sdfc_sim* s) {
    if (!s) return SDFC_OK;
    cudaSetDevice(s->ctx->device);
    cudaFree(s->d_bodies); cudaFree(s->d_pair_a); cudaFree(s->d_pair_b); cudaFree(s->d_state); cudaFree(s->d_scalars);
    cudaFree(s->d_ints); cudaFree(s->d_converged); cudaFree(s->d_touched); cudaFree(s->d_active); cudaFree(s->d_touched_ground); cudaFree(s->d_ground_body); cudaFree(s->d_fext); cudaFree(s->d_unfinished);
    cudaFree(s->d_counters); cudaFree(s->d_traj_q); cudaFree(s->d_traj_qd);
    if (s->h_unfinished) cudaFreeHost(s->h_unfinished);
    for (int l = 0; l < sdfc_sim::kMaxLanes; ++l) {
        if (s->lane_stream[l]) cudaStreamDestroy(s->lane_stream[l]);
        if (s->lane_event[l]) cudaEventDestroy(s->lane_event[l]);
    }
    if (s->fork_event) cudaEventDestroy(s->fork_event);
    delete s;
    return SDFC_OK;
}

int sdfc_sim_ndof_r(const sdfc_sim* s) { return s ? s->nr : 0; }

sdfc_status sdfc_sim_set_state(sdfc_sim* s, const double* q_host, const double* qdot_host) {
    SDFC_REQUIRE(s && q_host && qdot_host, "sdfc_sim_set_state: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    const size_t bytes = sizeof(double) * (size_t)s->n_envs * s->nr;
    SDFC_CUDA(cudaMemcpyAsync(s->dev.q, q_host, bytes, cudaMemcpyHostToDevice, s->ctx->stream));
    SDFC_CUDA(cudaMemcpyAsync(s->dev.qdot, qdot_host, bytes, cudaMemcpyHostToDevice, s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_state(sdfc_sim* s, double* q_host, double* qdot_host) {
    SDFC_REQUIRE(s && (q_host || qdot_host), "sdfc_sim_get_state: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    const size_t bytes = sizeof(double) * (size_t)s->n_envs * s->nr;
    if (q_host) SDFC_CUDA(cudaMemcpyAsync(q_host, s->dev.q, bytes, cudaMemcpyDeviceToHost, s->ctx->stream));
    if (qdot_host) SDFC_CUDA(cudaMemcpyAsync(qdot_host, s->dev.qdot, bytes, cudaMemcpyDeviceToHost, s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_set_external_force(sdfc_sim* s, const double* f_host) {
    SDFC_REQUIRE(s && f_host, "sdfc_sim_set_external_force: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    SDFC_CUDA(cudaMemcpyAsync(s->d_fext, f_host, sizeof(double) * (size_t)s->n_envs * s->n_bodies * 6, cudaMemcpyHostToDevice,
                              s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_clear_contact_history(sdfc_sim* s) {
    SDFC_REQUIRE(s != nullptr, "sdfc_sim_clear_contact_history: sim is NULL");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    SDFC_CUDA(cudaMemsetAsync(s->d_touched, 0, (size_t)s->n_envs * s->n_pairs, s->ctx->stream));
    if (s->n_grounds)
        SDFC_CUDA(cudaMemsetAsync(s->d_touched_ground, 0, (size_t)s->n_envs * s->n_grounds, s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_set_active(sdfc_sim* s, const uint8_t* active_host) {
    SDFC_REQUIRE(s != nullptr, "sdfc_sim_set_active: sim is NULL");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    s->use_active = active_host != nullptr;
    if (active_host) {
        SDFC_CUDA(cudaMemcpyAsync(s->d_active, active_host, (size_t)s->n_envs, cudaMemcpyHostToDevice, s->ctx->stream));
        SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    }
    return SDFC_OK;
}

sdfc_status sdfc_sim_forward(sdfc_sim* s, int n_steps, double* q_traj_host, double* qdot_traj_host) {
    SDFC_REQUIRE(s != nullptr, "sdfc_sim_forward: sim is NULL");
    SDFC_REQUIRE(n_steps >= 0, "sdfc_sim_forward: negative step count");
    if (n_steps == 0) return SDFC_OK;
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    const size_t per_step = (size_t)s->n_envs * s->nr;
    const bool want_traj = q_traj_host || qdot_traj_host;
    if (want_traj && s->traj_cap < (size_t)n_steps) {
        cudaFree(s->d_traj_q); cudaFree(s->d_traj_qd);
        s->d_traj_q = s->d_traj_qd = nullptr; s->traj_cap = 0;
        sdfc_status a = dev_alloc(&s->d_traj_q, per_step * n_steps);
        if (a != SDFC_OK) return a;
        a = dev_alloc(&s->d_traj_qd, per_step * n_steps);
        if (a != SDFC_OK) return a;
        s->traj_cap = (size_t)n_steps;
    }
    SimDev D = s->dev;
    D.q_traj = q_traj_host ? s->d_traj_q : nullptr;
    D.qdot_traj = qdot_traj_host ? s->d_traj_qd : nullptr;
    D.active = s->use_active ? s->d_active : nullptr;
    // Lanes: each environment range advances through its own rounds on its own stream; the
    // host feeds the lanes round-robin, so the 32-CTA assembly / solve kernels of one lane run
    // under the contact kernels of another. Results do not depend on the lane count.
    const int L = s->n_lanes;
    cudaStream_t ctx_stream = s->ctx->stream;
    SDFC_CUDA(cudaMemsetAsync(s->d_unfinished, 0, sizeof(int) * sdfc_sim::kMaxLanes, st));
    SDFC_CUDA(cudaEventRecord(s->fork_event, st));
    int lane_env0[sdfc_sim::kMaxLanes], lane_cnt[sdfc_sim::kMaxLanes];
    bool lane_live[sdfc_sim::kMaxLanes];
    for (int l = 0; l < L; ++l) {
        lane_env0[l] = (int)((int64_t)s->n_envs * l / L);
        lane_cnt[l] = (int)((int64_t)s->n_envs * (l + 1) / L) - lane_env0[l];
        lane_live[l] = lane_cnt[l] > 0;
        if (!lane_live[l]) continue;
        cudaStream_t ls = (L == 1) ? st : s->lane_stream[l];
        if (L > 1) SDFC_CUDA(cudaStreamWaitEvent(ls, s->fork_event, 0));
        SimDev Dl = D;
        Dl.env0 = lane_env0[l]; Dl.env_count = lane_cnt[l]; Dl.n_unfinished = s->d_unfinished + l;
        k_sim_begin<<<(lane_cnt[l] + 31) / 32, 32, 0, ls>>>(Dl, n_steps);
        s->ctx->launches += 1;
    }
    // every step needs at least two rounds (one Newton evaluation, one line-search
    // trial), so the unfinished-counter is only polled every second round
    const int64_t guard = (int64_t)n_steps * (D.max_iter * (kMaxLS + 1) + 1);
    int64_t depth = 0;
    sdfc_status rc = SDFC_OK;
    // The unfinished-counter of a lane is read back with a lag: the copy enqueued after round r
    // is only waited for before round r + 2 is issued, so the device always has queued work while
    // the host looks at an older answer (a finished lane runs at most two empty rounds).
    bool lane_pending[sdfc_sim::kMaxLanes] = {false, false, false, false};
    for (int64_t round = 0; round < guard && rc == SDFC_OK; ++round) {
        bool any = false;
        for (int l = 0; l < L && rc == SDFC_OK; ++l) {
            if (!lane_live[l]) continue;
            cudaStream_t ls = (L == 1) ? st : s->lane_stream[l];
            if ((round & 1) && lane_pending[l]) {
                if (cudaEventSynchronize(s->lane_event[l]) != cudaSuccess) { rc = fail(SDFC_ERR_CUDA, "sdfc_sim_forward: lane sync failed"); break; }
                lane_pending[l] = false;
                if (s->h_unfinished[l] <= 0) { lane_live[l] = false; continue; }
            }
            any = true;
            SimDev Dl = D;
            Dl.env0 = lane_env0[l]; Dl.env_count = lane_cnt[l]; Dl.n_unfinished = s->d_unfinished + l;
            const int blocks = (lane_cnt[l] + 31) / 32;
            s->ctx->stream = ls;  // the batch launches (and their timing events) go to the lane's stream
            launch_pre(s, Dl, 0, blocks, ls);
            rc = sdfc_host::batch_run(s->batch, D.E, D.phi, 0, D.env_mode, /*scatter=*/false, nullptr, nullptr, nullptr,
                                      nullptr, lane_env0[l], lane_cnt[l]);
            s->ctx->stream = ctx_stream;
            if (rc != SDFC_OK) break;
            launch_post(s, Dl, blocks, ls);
            s->ctx->launches += 2;
            if (round & 1) {
                if (cudaMemcpyAsync(s->h_unfinished + l, s->d_unfinished + l, sizeof(int), cudaMemcpyDeviceToHost, ls) != cudaSuccess ||
                    cudaEventRecord(s->lane_event[l], ls) != cudaSuccess)
                    rc = fail(SDFC_ERR_CUDA, "sdfc_sim_forward: poll failed");
                lane_pending[l] = true;
            }
        }
        if (!any) break;
        depth += 1;
    }
    s->rounds += depth;
    if (rc != SDFC_OK) { cudaDeviceSynchronize(); return rc; }
    if (L > 1)
        for (int l = 0; l < L; ++l) {  // join: the context stream continues after every lane
            SDFC_CUDA(cudaEventRecord(s->lane_event[l], s->lane_stream[l]));
            SDFC_CUDA(cudaStreamWaitEvent(st, s->lane_event[l], 0));
        }
    SDFC_CUDA(cudaGetLastError());
    if (q_traj_host) SDFC_CUDA(cudaMemcpyAsync(q_traj_host, s->d_traj_q, sizeof(double) * per_step * n_steps, cudaMemcpyDeviceToHost, st));
    if (qdot_traj_host) SDFC_CUDA(cudaMemcpyAsync(qdot_traj_host, s->d_traj_qd, sizeof(double) * per_step * n_steps, cudaMemcpyDeviceToHost, st));
    SDFC_CUDA(cudaStreamSynchronize(st));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_status(sdfc_sim* s, uint8_t* converged, uint8_t* touched) {
    SDFC_REQUIRE(s != nullptr, "sdfc_sim_get_status: sim is NULL");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    if (converged) SDFC_CUDA(cudaMemcpyAsync(converged, s->d_converged, s->n_envs, cudaMemcpyDeviceToHost, s->ctx->stream));
    if (touched) SDFC_CUDA(cudaMemcpyAsync(touched, s->d_touched, (size_t)s->n_envs * s->n_pairs, cudaMemcpyDeviceToHost, s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_ground_history(sdfc_sim* s, uint8_t* touched_ground) {
    SDFC_REQUIRE(s && touched_ground, "sdfc_sim_get_ground_history: NULL argument");
    SDFC_REQUIRE(s->n_grounds > 0, "sdfc_sim_get_ground_history: the simulation has no ground contacts");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    SDFC_CUDA(cudaMemcpyAsync(touched_ground, s->d_touched_ground, (size_t)s->n_envs * s->n_grounds, cudaMemcpyDeviceToHost,
                              s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_poses(sdfc_sim* s, double* E_host) {
    SDFC_REQUIRE(s && E_host, "sdfc_sim_get_poses: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    launch_pre(s, s->dev, 1, (s->n_envs + 31) / 32, s->ctx->stream);
    s->ctx->launches += 1;
    SDFC_CUDA(cudaMemcpyAsync(E_host, s->dev.E, sizeof(double) * (size_t)s->n_envs * s->n_bodies * 16, cudaMemcpyDeviceToHost,
                              s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_stats(sdfc_sim* s, int64_t stats[4]) {
    SDFC_REQUIRE(s && stats, "sdfc_sim_get_stats: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    unsigned long long c[4] = {0, 0, 0, 0};
    SDFC_CUDA(cudaMemcpyAsync(c, s->d_counters, sizeof(c), cudaMemcpyDeviceToHost, s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    stats[0] = s->rounds; stats[1] = (int64_t)c[0]; stats[2] = (int64_t)c[1]; stats[3] = (int64_t)c[2];
    return SDFC_OK;
}

}  // extern "C"
