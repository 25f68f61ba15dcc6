// sdfc_step_math.cuh — the per-environment arithmetic of the batched BDF1 stepper, written for
// ONE TEAM OF THREADS PER ENVIRONMENT (a warp; a whole CTA when few environments are left and the latency of
// one environment is all there is) with the environment's 6 x n / 12 x 12 blocks in SHARED MEMORY.
//
// Every routine is a sequence of data-parallel statements WFOR(e, N) { one output element } separated by
// WSYNC(): the threads of the team split the N independent outputs of a statement (an element of a 3 x 3
// product, a column of a Jacobian, an entry of the reduced matrix ...), all operands live in the warp's
// workspace, and nothing is kept in per-thread local memory. The per-environment work is 10^3 - 10^4
// dependent flops (exponential-map kinematics with second derivatives, reduced assembly, pivoted LU):
// one thread per environment runs them as one serial chain out of 11-19 KB local-memory frames (round 1);
// here a statement costs one shared-memory round trip regardless of N <= 32.
//
// The file also compiles for the HOST (tests/host_step_harness.cpp): WFOR then runs its elements one
// after the other, forwards or backwards, so the CPU suite checks the restated algebra against the numpy
// oracle and proves that no statement reads what the same statement writes (both orders give the same bits).
//
// Restated from (paths relative to simulation/redmax/ of the reference): Joint/Joint.cpp:95-141,
// JointFixed.cpp:6-13, JointTranslational.cpp:9-40, JointFree3DExp.cpp:13-110,
// JointSphericalExp.cpp:13-183,237-264, Body/Body.cpp:138-181,250-330, Robot.cpp:656-744,
// Simulation.cpp:379-503 (computeMatrices), :1081-1112 (BDF1 residual), :1028 (PartialPivLU).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define SDFC_HD __host__ __device__ __forceinline__
#else
#define SDFC_HD inline
#endif

namespace sdfc {

struct WarpExec {
    int lane;       // device: index of the thread in its team; host: unused
    int dir;        // host: +1 forward / -1 backward element order
    int nth = 32;   // device: threads of the team (32 = a warp; otherwise the whole CTA)
};

#if defined(__CUDA_ARCH__)
#define WFOR(e, N) for (int e = wx.lane; e < (N); e += wx.nth)
#define WSYNC() do { if (wx.nth > 32) __syncthreads(); else __syncwarp(); } while (0)
#define WLANE0 (wx.lane == 0)
#else
#define WFOR(e, N) for (int e##_i = 0, e = (wx.dir > 0 ? 0 : (N)-1); e##_i < (N); ++e##_i, e += wx.dir)
#define WSYNC() ((void)0)
#define WLANE0 true
#endif

constexpr int kJointFixed = 0, kJointTranslational = 1, kJointFree3DExp = 2;   // == SDFC_JOINT_* of the C-ABI
constexpr int kStepCapR = 12;    // reduced dofs per environment
constexpr int kStepCapDyn = 2;   // non-fixed bodies per environment
constexpr double kStepMathEps = 1e-9;  // math::eps, Utils.h:14

struct BodyDev {
    int type, dof_off, ndof, dyn;  // dyn: slot among non-fixed bodies, or -1
    double Rp[9], pp[3];           // Joint::_E_pj_0 (row-major rotation)
    double Rji[9], pji[3];         // Body::_E_ji
    double Aij[36];                // Ad(E_ij), row-major
    double I6[6], mass;
};

// ------------------------------------------------------------------ element helpers (row-major 3x3)
SDFC_HD double mul33_e(const double* A, const double* B, int r, int c) { return A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c]; }
SDFC_HD double tmul33_e(const double* A, const double* B, int r, int c) { return A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c]; }  // (A^T B)(r,c)
SDFC_HD double skew_e(const double* v, int i) {   // element i of math::skew(v), Utils.h:25-33
    return i == 1 ? -v[2] : i == 2 ? v[1] : i == 3 ? v[2] : i == 5 ? -v[0] : i == 6 ? -v[1] : i == 7 ? v[0] : 0.0;
}
// unskew(A) = (A[7], A[2], A[3]) = entries (2,1), (0,2), (1,0)
SDFC_HD int unskew_r(int a) { return a == 0 ? 2 : (a == 1 ? 0 : 1); }
SDFC_HD int unskew_c(int a) { return a == 0 ? 1 : (a == 1 ? 2 : 0); }
SDFC_HD double dot3h(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ------------------------------------------------------------------ serial pieces (rare paths, one lane)
SDFC_HD void s_skew(const double* v, double* S) {
    S[0] = 0; S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2]; S[4] = 0; S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
// math::exp(Vector3), Utils.h:111-127
SDFC_HD void s_exp3(const double* w, double* R) {
    const double wlen = sqrt(dot3h(w, w));
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (wlen > 1e-12) {
        const double wx_ = w[0] / wlen, wy = w[1] / wlen, wz = w[2] / wlen;
        const double c = cos(wlen), s = sin(wlen), c1 = 1.0 - c;
        R[0] = c + wx_ * wx_ * c1;      R[1] = -wz * s + wx_ * wy * c1; R[2] = wy * s + wx_ * wz * c1;
        R[3] = wz * s + wx_ * wy * c1;  R[4] = c + wy * wy * c1;        R[5] = -wx_ * s + wy * wz * c1;
        R[6] = -wy * s + wx_ * wz * c1; R[7] = wx_ * s + wy * wz * c1;  R[8] = c + wz * wz * c1;
    }
}

// JointSphericalExp::update (JointSphericalExp.cpp:13-183). T = _S_j top 3x3, Tdot = _S_j_dot top 3x3.
struct SphW {
    double R[9], Rdot[9], T[9], Tdot[9];
    double dR[3][9], dRdot[3][9], dT[3][9], dTdot[3][9];
    // scratch
    double Sq[9], Sqd[9], ImR[9];
    double A[3][9], BC[3][9], Adot[3][9], BCdot[3][9], P[3][9], Qm[3][9];   // P = Adot R, Qm = A Rdot
    double w1[3][3], w2[3][3];
    double w3[3][3][3];                      // [j][k][r]
    double dBC[3][3][9], dA[3][3][9], dAd[3][3][9];   // [j][k]
    double X1[3][3][9], X2[3][3][9], X5[3][3][9], X6[3][3][9], X8[3][3][9], X9[3][3][9];
};

// case |q| < eps (JointSphericalExp.cpp:27-66): closed form, evaluated by one lane
SDFC_HD void s_spherical_exp_small(const double* qd, SphW& o) {
    const double e[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 9; ++i) { o.R[i] = (i % 4 == 0) ? 1.0 : 0.0; o.Rdot[i] = 0; o.T[i] = o.R[i]; }
    for (int j = 0; j < 3; ++j) {
        s_skew(e[j], o.dR[j]);
        for (int i = 0; i < 9; ++i) { o.Rdot[i] += o.dR[j][i] * qd[j]; o.dRdot[j][i] = 0; o.dTdot[j][i] = 0; }
    }
    for (int j = 0; j < 3; ++j) {
        double sj[9];
        s_skew(e[j], sj);
        for (int a = 0; a < 3; ++a) o.Tdot[a * 3 + j] = tmul33_e(o.Rdot, sj, unskew_r(a), unskew_c(a));
        for (int k = 0; k < 3; ++k) {
            double sk[9];
            s_skew(e[k], sk);
            for (int a = 0; a < 3; ++a) o.dT[k][a * 3 + j] = tmul33_e(sk, sj, unskew_r(a), unskew_c(a));
        }
    }
}

// q, qd: the 3 rotational coordinates / rates (workspace memory). All lanes call this together.
SDFC_HD void w_spherical_exp(const WarpExec& wx, const double* q, const double* qd, bool deriv, SphW& o) {
    const double q2 = dot3h(q, q);
    if (sqrt(q2) < kStepMathEps) {
        if (WLANE0) s_spherical_exp_small(qd, o);
        WSYNC();
        return;
    }
    const double d = 1.0 / q2;
    const double qqd = dot3h(q, qd);
    const double ddot = -2.0 * qqd / (q2 * q2);
    if (WLANE0) {
        s_exp3(q, o.R);
        s_skew(q, o.Sq); s_skew(qd, o.Sqd);
        for (int i = 0; i < 9; ++i) o.ImR[i] = ((i % 4 == 0) ? 1.0 : 0.0) - o.R[i];
    }
    WSYNC();
    WFOR(e, 9) {   // w1[j] = Sq * col(ImR, j)
        const int j = e / 3, r = e % 3;
        o.w1[j][r] = o.Sq[r * 3] * o.ImR[j] + o.Sq[r * 3 + 1] * o.ImR[3 + j] + o.Sq[r * 3 + 2] * o.ImR[6 + j];
    }
    WSYNC();
    WFOR(e, 27) {
        const int j = e / 9, i = e % 9;
        const double bc = q[j] * o.Sq[i] + skew_e(o.w1[j], i);
        o.BC[j][i] = bc;
        o.A[j][i] = bc * d;
    }
    WSYNC();
    WFOR(e, 27) { const int j = e / 9, i = e % 9; o.dR[j][i] = mul33_e(o.A[j], o.R, i / 3, i % 3); }
    WSYNC();
    WFOR(i, 9) o.Rdot[i] = o.dR[0][i] * qd[0] + o.dR[1][i] * qd[1] + o.dR[2][i] * qd[2];
    WSYNC();
    WFOR(e, 9) {   // w2[j] = Sqd * col(ImR, j) - Sq * col(Rdot, j)
        const int j = e / 3, r = e % 3;
        const double a = o.Sqd[r * 3] * o.ImR[j] + o.Sqd[r * 3 + 1] * o.ImR[3 + j] + o.Sqd[r * 3 + 2] * o.ImR[6 + j];
        const double b = o.Sq[r * 3] * o.Rdot[j] + o.Sq[r * 3 + 1] * o.Rdot[3 + j] + o.Sq[r * 3 + 2] * o.Rdot[6 + j];
        o.w2[j][r] = a - b;
    }
    WSYNC();
    WFOR(e, 27) {
        const int j = e / 9, i = e % 9;
        const double bcd = qd[j] * o.Sq[i] + q[j] * o.Sqd[i] + skew_e(o.w2[j], i);
        o.BCdot[j][i] = bcd;
        o.Adot[j][i] = bcd * d + o.BC[j][i] * ddot;
    }
    WSYNC();
    WFOR(e, 54) {
        const int which = e / 27, j = (e % 27) / 9, i = e % 9;
        if (which == 0) o.P[j][i] = mul33_e(o.Adot[j], o.R, i / 3, i % 3);
        else o.Qm[j][i] = mul33_e(o.A[j], o.Rdot, i / 3, i % 3);
    }
    WSYNC();
    WFOR(e, 18) {
        const int which = e / 9, j = (e % 9) / 3, a = e % 3;
        const int r = unskew_r(a), c = unskew_c(a);
        if (which == 0) o.T[a * 3 + j] = tmul33_e(o.R, o.dR[j], r, c);
        else o.Tdot[a * 3 + j] = tmul33_e(o.Rdot, o.dR[j], r, c) + tmul33_e(o.R, o.P[j], r, c) + tmul33_e(o.R, o.Qm[j], r, c);
    }
    WSYNC();
    if (!deriv) return;
    const double q22 = q2 * q2;
    // w3[j][k] = skew(e_k) col(ImR, j) - Sq col(dR[k], j)
    WFOR(e, 27) {
        const int j = e / 9, k = (e % 9) / 3, r = e % 3;
        double ek[3] = {0, 0, 0};
        ek[k] = 1.0;
        const double a = skew_e(ek, r * 3) * o.ImR[j] + skew_e(ek, r * 3 + 1) * o.ImR[3 + j] + skew_e(ek, r * 3 + 2) * o.ImR[6 + j];
        const double b = o.Sq[r * 3] * o.dR[k][j] + o.Sq[r * 3 + 1] * o.dR[k][3 + j] + o.Sq[r * 3 + 2] * o.dR[k][6 + j];
        o.w3[j][k][r] = a - b;
    }
    WSYNC();
    WFOR(e, 81) {
        const int j = e / 27, k = (e % 27) / 9, i = e % 9;
        double ek[3] = {0, 0, 0};
        ek[k] = 1.0;
        double db = q[j] * skew_e(ek, i);
        if (j == k) db += o.Sq[i];
        const double dbc = db + skew_e(o.w3[j][k], i);
        o.dBC[j][k][i] = dbc;
        o.dA[j][k][i] = dbc * d + o.BC[j][i] * (-2.0 * q[k] / q22);
    }
    WSYNC();
    WFOR(e, 162) {   // X1 = dA[j][k] R, X2 = A[j] dR[k]  (d2R/dq_j dq_k = X1 + X2)
        const int which = e / 81, j = (e % 81) / 27, k = (e % 27) / 9, i = e % 9;
        if (which == 0) o.X1[j][k][i] = mul33_e(o.dA[j][k], o.R, i / 3, i % 3);
        else o.X2[j][k][i] = mul33_e(o.A[j], o.dR[k], i / 3, i % 3);
    }
    WSYNC();
    WFOR(e, 27) {
        const int k = e / 9, i = e % 9;
        o.dRdot[k][i] = (o.X1[0][k][i] + o.X2[0][k][i]) * qd[0] + (o.X1[1][k][i] + o.X2[1][k][i]) * qd[1] +
                        (o.X1[2][k][i] + o.X2[2][k][i]) * qd[2];
    }
    WSYNC();
    WFOR(e, 27) {   // w3[j][k] reused: -Sqd col(dR[k], j) - skew(e_k) col(Rdot, j) - Sq col(dRdot[k], j)
        const int j = e / 9, k = (e % 9) / 3, r = e % 3;
        double ek[3] = {0, 0, 0};
        ek[k] = 1.0;
        const double a = o.Sqd[r * 3] * o.dR[k][j] + o.Sqd[r * 3 + 1] * o.dR[k][3 + j] + o.Sqd[r * 3 + 2] * o.dR[k][6 + j];
        const double b = skew_e(ek, r * 3) * o.Rdot[j] + skew_e(ek, r * 3 + 1) * o.Rdot[3 + j] + skew_e(ek, r * 3 + 2) * o.Rdot[6 + j];
        const double c = o.Sq[r * 3] * o.dRdot[k][j] + o.Sq[r * 3 + 1] * o.dRdot[k][3 + j] + o.Sq[r * 3 + 2] * o.dRdot[k][6 + j];
        o.w3[j][k][r] = -a - b - c;
    }
    WSYNC();
    WFOR(e, 81) {
        const int j = e / 27, k = (e % 27) / 9, i = e % 9;
        double ek[3] = {0, 0, 0};
        ek[k] = 1.0;
        double dbd = qd[j] * skew_e(ek, i);
        if (j == k) dbd += o.Sqd[i];
        const double dd_dq = -2.0 * q[k] / q22;
        const double dddot_dq = (-2.0 * qd[k] * q2 + 8.0 * qqd * q[k]) / (q22 * q2);
        o.dAd[j][k][i] = (dbd + skew_e(o.w3[j][k], i)) * d + o.BCdot[j][i] * dd_dq + o.dBC[j][k][i] * ddot + o.BC[j][i] * dddot_dq;
    }
    WSYNC();
    WFOR(e, 324) {   // X5 = dAd R, X6 = Adot[j] dR[k], X8 = dA Rdot, X9 = A[j] dRdot[k]
        const int which = e / 81, j = (e % 81) / 27, k = (e % 27) / 9, i = e % 9;
        const int r = i / 3, c = i % 3;
        if (which == 0) o.X5[j][k][i] = mul33_e(o.dAd[j][k], o.R, r, c);
        else if (which == 1) o.X6[j][k][i] = mul33_e(o.Adot[j], o.dR[k], r, c);
        else if (which == 2) o.X8[j][k][i] = mul33_e(o.dA[j][k], o.Rdot, r, c);
        else o.X9[j][k][i] = mul33_e(o.A[j], o.dRdot[k], r, c);
    }
    WSYNC();
    WFOR(e, 54) {
        const int which = e / 27, j = (e % 27) / 9, k = (e % 9) / 3, a = e % 3;
        const int r = unskew_r(a), c = unskew_c(a);
        if (which == 0) {
            // dT[k].col(j) = unskew(dR_k^T dR_j + R^T d2R)
            o.dT[k][a * 3 + j] = tmul33_e(o.dR[k], o.dR[j], r, c) + (tmul33_e(o.R, o.X1[j][k], r, c) + tmul33_e(o.R, o.X2[j][k], r, c));
        } else {
            // dTdot[k].col(j): the nine products of JointSphericalExp.cpp:166-169, in that order
            double acc = tmul33_e(o.dRdot[k], o.dR[j], r, c);     // dRdot_k^T A_j R
            acc += tmul33_e(o.Rdot, o.X1[j][k], r, c);            // Rdot^T dA_jk R
            acc += tmul33_e(o.Rdot, o.X2[j][k], r, c);            // Rdot^T A_j dR_k
            acc += tmul33_e(o.dR[k], o.P[j], r, c);               // dR_k^T Adot_j R
            acc += tmul33_e(o.R, o.X5[j][k], r, c);               // R^T dAd_jk R
            acc += tmul33_e(o.R, o.X6[j][k], r, c);               // R^T Adot_j dR_k
            acc += tmul33_e(o.dR[k], o.Qm[j], r, c);              // dR_k^T A_j Rdot
            acc += tmul33_e(o.R, o.X8[j][k], r, c);               // R^T dA_jk Rdot
            acc += tmul33_e(o.R, o.X9[j][k], r, c);               // R^T A_j dRdot_k
            o.dTdot[k][a * 3 + j] = acc;
        }
    }
    WSYNC();
}

// Joint quantities of one non-fixed body: S_j, S_j_dot (6 x n, row stride 6) and their q-derivatives
// (nonzero only for the rotational dofs 3..5 of a free3d-exp joint), plus Q = [R p; 0 1].
struct JointW {
    double R[9], p[3];
    double S[36], Sd[36];
    double dS[3][36], dSd[3][36];
};

// q, qd: this body's slice of the candidate state. JointTranslational.cpp:9-40, JointFree3DExp.cpp:13-110.
SDFC_HD void w_joint_kin(const WarpExec& wx, const BodyDev& b, const double* q, const double* qd, bool deriv, JointW& k, SphW& sph) {
    const bool fr = b.type == kJointFree3DExp;
    if (fr) w_spherical_exp(wx, q + 3, qd + 3, deriv, sph);
    WFOR(e, 72) {
        const int which = e / 36, i = e % 36, r = i / 6, c = i % 6;
        double v = 0.0;
        if (b.type == kJointTranslational) {
            if (which == 0 && r >= 3 && c == r - 3) v = 1.0;
        } else if (fr) {
            if (r < 3 && c >= 3) v = which == 0 ? sph.T[r * 3 + c - 3] : sph.Tdot[r * 3 + c - 3];
            else if (r >= 3 && c < 3) v = which == 0 ? sph.R[c * 3 + r - 3] : sph.Rdot[c * 3 + r - 3];
        }
        if (which == 0) k.S[i] = v; else k.Sd[i] = v;
    }
    WFOR(i, 12) {
        if (i < 9) k.R[i] = fr ? sph.R[i] : ((i % 4 == 0) ? 1.0 : 0.0);
        else k.p[i - 9] = (b.type == kJointFixed) ? 0.0 : q[i - 9];
    }
    if (fr && deriv) {
        WFOR(e, 216) {
            const int which = e / 108, d = (e % 108) / 36, i = e % 36, r = i / 6, c = i % 6;
            double v = 0.0;
            if (r < 3 && c >= 3) v = which == 0 ? sph.dT[d][r * 3 + c - 3] : sph.dTdot[d][r * 3 + c - 3];
            else if (r >= 3 && c < 3) v = which == 0 ? sph.dR[d][c * 3 + r - 3] : sph.dRdot[d][c * 3 + r - 3];
            if (which == 0) k.dS[d][i] = v; else k.dSd[d][i] = v;
        }
    }
    WSYNC();
}

// E_0i = E_pj_0 * Q * E_ji (Joint.cpp:101-113, Body.cpp:140); phi_i = A_ij (S_j qdot) (Joint.cpp:119-121,
// Body.cpp:156). E16 (column-major 4x4), phi6, R0i may each be nullptr. tmp: 24 doubles of workspace.
SDFC_HD void w_body_pose(const WarpExec& wx, const BodyDev& b, const JointW& k, const double* qd, double* E16, double* phi6,
                         double* R0i, double* tmp) {
    double* R0j = tmp;        // 9
    double* p0j = tmp + 9;    // 3
    double* pj = tmp + 12;    // 6
    WFOR(e, 18) {
        if (e < 9) R0j[e] = mul33_e(b.Rp, k.R, e / 3, e % 3);
        else if (e < 12) { const int r = e - 9; p0j[r] = (b.Rp[r * 3] * k.p[0] + b.Rp[r * 3 + 1] * k.p[1] + b.Rp[r * 3 + 2] * k.p[2]) + b.pp[r]; }
        else {
            const int r = e - 12;
            double s = 0;
            for (int c = 0; c < b.ndof; ++c) s += k.S[r * 6 + c] * qd[c];
            pj[r] = s;
        }
    }
    WSYNC();
    WFOR(e, 18) {
        if (e < 9) {
            const double v = mul33_e(R0j, b.Rji, e / 3, e % 3);
            if (E16) E16[(e / 3) + 4 * (e % 3)] = v;
            if (R0i) R0i[e] = v;
        } else if (e < 12) {
            const int r = e - 9;
            const double v = (R0j[r * 3] * b.pji[0] + R0j[r * 3 + 1] * b.pji[1] + R0j[r * 3 + 2] * b.pji[2]) + p0j[r];
            if (E16) { E16[r + 12] = v; E16[3 + 4 * r] = 0.0; }
        } else {
            const int r = e - 12;
            double s = 0;
            for (int c = 0; c < 6; ++c) s += b.Aij[r * 6 + c] * pj[c];
            if (phi6) phi6[r] = (b.ndof > 0) ? s : 0.0;
        }
    }
    if (E16 && WLANE0) E16[15] = 1.0;
    WSYNC();
}

// Per non-fixed body: Jacobian blocks, maximal-coordinate force data and the reduced-assembly temporaries.
struct DynW {
    double J[36], Jd[36];            // 6 x n (row stride 6)
    double dJ[3][36], dJd[3][36];    // w.r.t. rotational dofs (free3d-exp only)
    double fm[6], R0i[9], phi[6];
    double MmJ[36], M[36], JTMmJd[36], MmJdq[6], dq[6], dJq[36], v6[3][6];
    int n, off, body, has_dJ;
};

struct EnvW {
    double xc[kStepCapR], qdc[kStepCapR];
    SphW sph;
    JointW jk;
    DynW dyn[kStepCapDyn];
    double Kmm[kStepCapDyn][kStepCapDyn][36], Dmm[kStepCapDyn][kStepCapDyn][36];   // maximal blocks between non-fixed bodies
    double g[kStepCapR], H[kStepCapR * kStepCapR], rhs[kStepCapR];
    double JTK[36], JTD[36];
    double tmp[48];
    int piv;
};

// Kinematics of non-fixed body b into its DynW + Body::computeMaximalForce[WithDerivative]'s Coriolis,
// gravity and external terms (Body.cpp:250-330). fext6: this body's external force input.
SDFC_HD void w_dyn_body(const WarpExec& wx, const BodyDev& B, int body, bool deriv, const double* grav, const double* fext6,
                        EnvW& W) {
    DynW& D = W.dyn[B.dyn];
    const double* qd = W.qdc + B.dof_off;
    if (WLANE0) { D.n = B.ndof; D.off = B.dof_off; D.body = body; D.has_dJ = (deriv && B.type == kJointFree3DExp) ? 1 : 0; }
    w_joint_kin(wx, B, W.xc + B.dof_off, qd, deriv, W.jk, W.sph);
    w_body_pose(wx, B, W.jk, qd, nullptr, D.phi, D.R0i, W.tmp);
    const bool has_dJ = deriv && B.type == kJointFree3DExp;
    const int n = B.ndof;
    // J = A_ij S etc. (Robot.cpp:664-665)
    WFOR(e, (has_dJ ? 8 : 2) * 36) {
        const int which = e / 36, i = e % 36, r = i / 6, c = i % 6;
        const double* src = which == 0 ? W.jk.S : which == 1 ? W.jk.Sd : which < 5 ? W.jk.dS[which - 2] : W.jk.dSd[which - 5];
        double s = 0;
        if (c < n)
            for (int k = 0; k < 6; ++k) s += B.Aij[r * 6 + k] * src[k * 6 + c];
        double* dst = which == 0 ? D.J : which == 1 ? D.Jd : which < 5 ? D.dJ[which - 2] : D.dJd[which - 5];
        dst[i] = s;
    }
    // body forces: one lane (about 100 flops), then the 6x6 blocks in parallel
    double* t = W.tmp;   // fg[3] e1[3] e2[3] Iw[3] mv[3]
    if (WLANE0) {
        const double* w = D.phi; const double* v = D.phi + 3;
        double Iw[3], mv[3], c1[3], c2[3], c3[3];
        for (int a = 0; a < 3; ++a) { Iw[a] = B.I6[a] * w[a]; mv[a] = B.I6[3 + a] * v[a]; }
        c1[0] = w[1] * Iw[2] - w[2] * Iw[1]; c1[1] = w[2] * Iw[0] - w[0] * Iw[2]; c1[2] = w[0] * Iw[1] - w[1] * Iw[0];
        c2[0] = v[1] * mv[2] - v[2] * mv[1]; c2[1] = v[2] * mv[0] - v[0] * mv[2]; c2[2] = v[0] * mv[1] - v[1] * mv[0];
        c3[0] = w[1] * mv[2] - w[2] * mv[1]; c3[1] = w[2] * mv[0] - w[0] * mv[2]; c3[2] = w[0] * mv[1] - w[1] * mv[0];
        const double ge[3] = {B.mass * grav[0], B.mass * grav[1], B.mass * grav[2]};
        for (int r = 0; r < 3; ++r) {   // R^T v
            t[r] = D.R0i[r] * ge[0] + D.R0i[3 + r] * ge[1] + D.R0i[6 + r] * ge[2];
            t[3 + r] = D.R0i[r] * fext6[0] + D.R0i[3 + r] * fext6[1] + D.R0i[6 + r] * fext6[2];
            t[6 + r] = D.R0i[r] * fext6[3] + D.R0i[3 + r] * fext6[4] + D.R0i[6 + r] * fext6[5];
            t[9 + r] = Iw[r]; t[12 + r] = mv[r];
        }
        for (int a = 0; a < 3; ++a) {
            D.fm[a] = (-c1[a] - c2[a]) + B.I6[a] * t[3 + a];
            D.fm[3 + a] = -c3[a] + t[a] + B.I6[3 + a] * t[6 + a];
        }
    }
    WSYNC();
    if (deriv) {
        double* K = W.Kmm[B.dyn][B.dyn];
        double* Dm = W.Dmm[B.dyn][B.dyn];
        const double* w = D.phi; const double* v = D.phi + 3;
        WFOR(e, 9) {
            const int r = e / 3, c = e % 3;
            K[(3 + r) * 6 + c] += skew_e(t, e) + skew_e(t + 6, e) * B.I6[3 + r];
            K[r * 6 + c] += skew_e(t + 3, e) * B.I6[r];
            // Dm = ad^T diag(I) - [ [Iw] [mv]; [mv] 0 ]   (Body.cpp:305-329)
            Dm[r * 6 + c] += skew_e(w, c * 3 + r) * B.I6[c] + skew_e(t + 9, e);
            Dm[r * 6 + 3 + c] += skew_e(v, c * 3 + r) * B.I6[3 + c] + skew_e(t + 12, e);
            Dm[(3 + r) * 6 + 3 + c] += skew_e(w, c * 3 + r) * B.I6[3 + c];
            Dm[(3 + r) * 6 + c] += skew_e(t + 12, e);
        }
    }
    WSYNC();
}

// Reduced assembly (Simulation.cpp:379-503) and the BDF1 residual / Jacobian (:1081-1112):
// g = M dq - h^2 f, H = M - h^2 K - h D, from the DynW blocks and the maximal Kmm / Dmm.
// q0, qdot0: the environment's previous state (global memory).
SDFC_HD void w_assemble(const WarpExec& wx, const BodyDev* bodies, int n_dyn, int nr, bool deriv, double h, const double* q0,
                        const double* qdot0, EnvW& W) {
    constexpr int LD = kStepCapR;
    WFOR(e, nr * (deriv ? nr + 1 : 1)) {
        if (e < nr) W.g[e] = 0.0; else W.H[((e - nr) / nr) * LD + (e - nr) % nr] = 0.0;
    }
    WSYNC();
    for (int s = 0; s < n_dyn; ++s) {
        DynW& D = W.dyn[s];
        const BodyDev& B = bodies[D.body];
        const int n = D.n, off = D.off;
        const double* qd = W.qdc + off;
        WFOR(e, 42 + n) {
            if (e < 36) { const int r = e / 6; D.MmJ[e] = D.J[e] * B.I6[r]; }
            else if (e < 42) {
                const int r = e - 36;
                double s2 = 0;
                for (int c = 0; c < n; ++c) s2 += D.Jd[r * 6 + c] * B.I6[r] * qd[c];
                D.MmJdq[r] = s2;
            } else {
                const int i = e - 42;
                D.dq[i] = W.xc[off + i] - q0[off + i] - h * qdot0[off + i];
            }
        }
        WSYNC();
        WFOR(e, n * n) {
            const int i = e / n, j = e % n;
            double m = 0, jj = 0;
            for (int r = 0; r < 6; ++r) { m += D.J[r * 6 + i] * D.MmJ[r * 6 + j]; jj += D.MmJ[r * 6 + i] * D.Jd[r * 6 + j]; }
            D.M[i * 6 + j] = m;
            D.JTMmJd[i * 6 + j] = jj;
        }
        if (deriv) {
            WFOR(e, 36 + 18) {
                if (e < 36) {   // dJq(:,k) = dJ_dq(k) qdot
                    const int r = e / 6, k = e % 6;
                    double s2 = 0;
                    if (D.has_dJ && k >= 3 && k < n)
                        for (int c = 0; c < n; ++c) s2 += D.dJ[k - 3][r * 6 + c] * qd[c];
                    D.dJq[r * 6 + k] = s2;
                } else if (D.has_dJ) {   // v6[k] = dJdot_dq(k) qdot
                    const int k = (e - 36) / 6, r = (e - 36) % 6;
                    double s2 = 0;
                    for (int c = 0; c < n; ++c) s2 += D.dJd[k][r * 6 + c] * qd[c];
                    D.v6[k][r] = s2;
                }
            }
        }
        WSYNC();
        // f = J^T fm - J^T Mm Jdot qdot ; g = M dq - h^2 f
        WFOR(i, n) {
            double f = 0, md = 0;
            for (int r = 0; r < 6; ++r) f += D.J[r * 6 + i] * (D.fm[r] - D.MmJdq[r]);
            for (int j = 0; j < n; ++j) md += D.M[i * 6 + j] * D.dq[j];
            W.g[off + i] = md - h * h * f;
        }
        if (deriv) {
            WFOR(e, n * n) {
                const int i = e / n, j = e % n;
                // H = M - h^2 K - h D : the M, Kqvv and Dqvv parts of the diagonal block
                double t = 0;
                for (int r = 0; r < 6; ++r) t += D.MmJ[r * 6 + i] * D.dJq[r * 6 + j];
                const double Dqvv = -D.JTMmJd[i * 6 + j] - t;
                double hv = D.M[i * 6 + j] - h * Dqvv;
                if (D.has_dJ && j >= 3) {
                    const int k = j;
                    const double* dJk = D.dJ[k - 3];
                    // Kqvv(:,k) = -dJ_k^T MmJdq - JTMm v6 ; plus dJ_k^T fm ; dM_dq(k) dq
                    double a = 0, b2 = 0, c2 = 0, dm = 0;
                    for (int r = 0; r < 6; ++r) {
                        a += dJk[r * 6 + i] * D.MmJdq[r];
                        b2 += D.MmJ[r * 6 + i] * D.v6[k - 3][r];
                        c2 += dJk[r * 6 + i] * D.fm[r];
                    }
                    for (int jj = 0; jj < n; ++jj) {
                        double t1 = 0, t2 = 0;
                        for (int r = 0; r < 6; ++r) { t1 += dJk[r * 6 + i] * D.MmJ[r * 6 + jj]; t2 += dJk[r * 6 + jj] * D.MmJ[r * 6 + i]; }
                        dm += (t1 + t2) * D.dq[jj];
                    }
                    hv += -h * h * ((-a - b2) + c2) + dm;
                }
                W.H[(off + i) * LD + off + j] += hv;
            }
        }
        WSYNC();
    }
    if (deriv) {
        // cross terms: J_s^T Km_st J_t, J_s^T Dm_st dJq_t (into K) and J_s^T Dm_st J_t (into D)
        for (int s = 0; s < n_dyn; ++s)
            for (int t = 0; t < n_dyn; ++t) {
                const DynW& A = W.dyn[s];
                const DynW& Bt = W.dyn[t];
                WFOR(e, A.n * 6) {
                    const int i = e / 6, c = e % 6;
                    double a = 0, b2 = 0;
                    for (int r = 0; r < 6; ++r) { a += A.J[r * 6 + i] * W.Kmm[s][t][r * 6 + c]; b2 += A.J[r * 6 + i] * W.Dmm[s][t][r * 6 + c]; }
                    W.JTK[i * 6 + c] = a; W.JTD[i * 6 + c] = b2;
                }
                WSYNC();
                WFOR(e, A.n * Bt.n) {
                    const int i = e / Bt.n, j = e % Bt.n;
                    double k1 = 0, k2 = 0, d1 = 0;
                    for (int c = 0; c < 6; ++c) {
                        k1 += W.JTK[i * 6 + c] * Bt.J[c * 6 + j];
                        k2 += W.JTD[i * 6 + c] * Bt.dJq[c * 6 + j];
                        d1 += W.JTD[i * 6 + c] * Bt.J[c * 6 + j];
                    }
                    W.H[(A.off + i) * LD + Bt.off + j] += -h * h * (k1 + k2) - h * d1;
                }
                WSYNC();
            }
    }
}

// Solve H x = rhs in place (n <= 12, row stride kStepCapR) by LU with partial pivoting, as Eigen's
// PartialPivLU (Simulation.cpp:1028): largest |entry| of the column, first one on ties.
SDFC_HD void w_lu_solve(const WarpExec& wx, double* A, double* b, int n, int* piv_slot) {
    constexpr int LD = kStepCapR;
    for (int c = 0; c < n; ++c) {
        if (WLANE0) {
            int piv = c;
            double best = fabs(A[c * LD + c]);
            for (int r = c + 1; r < n; ++r) {
                const double v = fabs(A[r * LD + c]);
                if (v > best) { best = v; piv = r; }
            }
            *piv_slot = piv;
        }
        WSYNC();
        const int piv = *piv_slot;
        if (piv != c) {
            WFOR(k, n + 1) {
                if (k < n) { const double t = A[c * LD + k]; A[c * LD + k] = A[piv * LD + k]; A[piv * LD + k] = t; }
                else { const double t = b[c]; b[c] = b[piv]; b[piv] = t; }
            }
            WSYNC();
        }
        const double inv = 1.0 / A[c * LD + c];
        WFOR(r0, n - c - 1) { const int r = c + 1 + r0; A[r * LD + c] = A[r * LD + c] * inv; }
        WSYNC();
        const int m = n - c - 1;
        WFOR(e, m * (m + 1)) {
            const int r = c + 1 + e / (m + 1), kk = e % (m + 1);
            const double f = A[r * LD + c];
            if (kk < m) A[r * LD + c + 1 + kk] -= f * A[c * LD + c + 1 + kk];
            else b[r] -= f * b[c];
        }
        WSYNC();
    }
    for (int r = n - 1; r >= 0; --r) {   // column-oriented back substitution
        if (WLANE0) b[r] = b[r] / A[r * LD + r];
        WSYNC();
        WFOR(i, r) b[i] -= A[i * LD + r] * b[r];
        WSYNC();
    }
}

// Everything one evaluation of an environment reads besides its workspace (device pointers in the
// kernel, plain arrays in the host harness). Per-environment pointers are already offset to the environment.
struct EvalIn {
    const BodyDev* bodies;
    int n_bodies, n_pairs, n_dyn, nr, n_grounds;
    const int *pair_a, *pair_b;        // body slots of every SDF contact force, in force order
    const int* ground_body;
    double h, grav[3];
    const double *q0, *qdot0;          // previous state [nr]
    const double* fext;                // [n_bodies][6]
    const double *fm_sub, *Km_sub, *Dm_sub;   // [n_pairs][12], [n_pairs][144] column-major
    const double* g_out;               // [n_grounds][78]: fm 6, Km 36, Dm 36 (column-major)
};

// One evaluation of the BDF1 residual (and its Jacobian) at the candidate state W.xc / W.qdc:
// Robot::update -> computeMatrices -> evaluate_g[_with_derivatives] (Simulation.cpp:379-503,1081-1112).
// Result: W.g (and W.H, row stride kStepCapR).
SDFC_HD void w_evaluate(const WarpExec& wx, const EvalIn& in, bool deriv, EnvW& W) {
    if (deriv) {
        WFOR(e, kStepCapDyn * kStepCapDyn * 36) { (&W.Kmm[0][0][0])[e] = 0.0; (&W.Dmm[0][0][0])[e] = 0.0; }
        WSYNC();
    }
    for (int b = 0; b < in.n_bodies; ++b) {
        const BodyDev& B = in.bodies[b];
        if (B.dyn < 0) continue;
        w_dyn_body(wx, B, b, deriv, in.grav, in.fext + 6 * b, W);
    }
    // contact blocks, in force order (Robot.cpp:524-525, 546-547); ground contacts first (XML order)
    for (int gi = 0; gi < in.n_grounds; ++gi) {
        const int sl = in.bodies[in.ground_body[gi]].dyn;
        if (sl < 0) continue;
        const double* o = in.g_out + (size_t)gi * 78;
        WFOR(e, deriv ? 42 : 6) {
            if (e < 6) W.dyn[sl].fm[e] += o[e];
            else {
                const int i = e - 6, r = i / 6, c = i % 6;   // body-diagonal 6x6 blocks (ForceGroundContact.cpp:129,186-189)
                W.Kmm[sl][sl][i] += o[6 + r + 6 * c];
                W.Dmm[sl][sl][i] += o[42 + r + 6 * c];
            }
        }
        WSYNC();
    }
    for (int p = 0; p < in.n_pairs; ++p) {
        const int sl[2] = {in.bodies[in.pair_a[p]].dyn, in.bodies[in.pair_b[p]].dyn};
        WFOR(e, deriv ? 156 : 12) {
            if (e < 12) {
                const int br = e / 6, r = e % 6;
                if (sl[br] >= 0) W.dyn[sl[br]].fm[r] += in.fm_sub[(size_t)p * 12 + 6 * br + r];
            } else {
                const int i = e - 12, blk = i / 36, r = (i % 36) / 6, c = i % 6, br = blk >> 1, bc = blk & 1;
                if (sl[br] >= 0 && sl[bc] >= 0) {
                    const size_t src = (size_t)p * 144 + (6 * br + r) + 12 * (6 * bc + c);   // column-major 12x12
                    W.Kmm[sl[br]][sl[bc]][r * 6 + c] += in.Km_sub[src];
                    W.Dmm[sl[br]][sl[bc]][r * 6 + c] += in.Dm_sub[src];
                }
            }
        }
        WSYNC();
    }
    w_assemble(wx, in.bodies, in.n_dyn, in.nr, deriv, in.h, in.q0, in.qdot0, W);
}

// Body poses E_0i (column-major 4x4) and twists phi of EVERY body at the candidate state W.xc / W.qdc:
// the contact path's inputs for the next evaluation (Robot::update, Joint.cpp:95-141, Body.cpp:138-181).
SDFC_HD void w_poses(const WarpExec& wx, const BodyDev* bodies, int n_bodies, EnvW& W, double* E, double* phi) {
    for (int b = 0; b < n_bodies; ++b) {
        const BodyDev& B = bodies[b];
        w_joint_kin(wx, B, W.xc + B.dof_off, W.qdc + B.dof_off, false, W.jk, W.sph);
        w_body_pose(wx, B, W.jk, W.qdc + B.dof_off, E + 16 * b, phi + 6 * b, nullptr, W.tmp);
    }
}

// JointSphericalExp::reparam (JointSphericalExp.cpp:237-264) of the rotational coordinates r / rates rd
// (workspace memory) when |r| > 1.5 pi: r <- r (1 - 2 pi / |r|) until it is not, rd from S_new rd_new = S_old rd_old.
SDFC_HD void w_reparam(const WarpExec& wx, double* r, double* rd, EnvW& W) {
    const double pi = acos(-1.0);
    if (!(sqrt(dot3h(r, r)) > 1.5 * pi)) return;
    w_spherical_exp(wx, r, rd, false, W.sph);
    WFOR(a, 3) W.tmp[a] = W.sph.T[a * 3] * rd[0] + W.sph.T[a * 3 + 1] * rd[1] + W.sph.T[a * 3 + 2] * rd[2];
    WSYNC();
    if (WLANE0) {
        double qn = sqrt(dot3h(r, r));
        while (qn > 1.5 * pi) {
            const double f = 1.0 - 2.0 * pi / qn;
            for (int a = 0; a < 3; ++a) r[a] *= f;
            qn = sqrt(dot3h(r, r));
        }
    }
    WSYNC();
    w_spherical_exp(wx, r, rd, false, W.sph);
    WFOR(e, 12) {
        if (e < 9) W.H[(e / 3) * kStepCapR + e % 3] = W.sph.T[e];
        else W.rhs[e - 9] = W.tmp[e - 9];
    }
    WSYNC();
    w_lu_solve(wx, W.H, W.rhs, 3, &W.piv);
    WFOR(a, 3) rd[a] = W.rhs[a];
    WSYNC();
}

// BodyDev from the C-ABI's body description (column-major 4x4 E_pj_0, E_ji; Utils.h:76-97 for Ad(E_ij)).
inline void make_body_dev(int joint_type, const double* E_pj_0, const double* E_ji, const double* inertia6, double mass,
                          int dof_off, int dyn_slot, BodyDev& B) {
    B.type = joint_type;
    B.ndof = joint_type == kJointFixed ? 0 : (joint_type == kJointTranslational ? 3 : 6);
    B.dof_off = dof_off;
    B.dyn = B.ndof > 0 ? dyn_slot : -1;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) { B.Rp[r * 3 + c] = E_pj_0[r + 4 * c]; B.Rji[r * 3 + c] = E_ji[r + 4 * c]; }
        B.pp[r] = E_pj_0[r + 12];
        B.pji[r] = E_ji[r + 12];
    }
    // A_ij = Ad(E_ij), E_ij = Einv(E_ji): R = Rji^T, p = -Rji^T pji
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
    for (int i = 0; i < 6; ++i) B.I6[i] = inertia6[i];
    B.mass = mass;
}

}  // namespace sdfc
