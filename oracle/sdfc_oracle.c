/* sdfc_oracle.c — CPU restatement of the reference's SDF-contact hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (the CUDA library, the
 * C-ABI, the host mirror) may include, link or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker.
 *
 * Parity status: PINNED. This restatement is validated (tests/test_oracle_*.py,
 * oracle/make_golden.py) against outputs of the unmodified reference compiled
 * into oracle/_ref (ref_probe), and against the committed golden vectors in
 * tests/golden/ that the same reference produced.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/simulation/redmax). Plain scalar C, fp64, evaluation order
 * chosen to match what Eigen 3.3.4 + g++ -O3 (no FMA) emit for the
 * bit-critical part (transforms, cell index, trilinear sample, FD stencils);
 * compile with -ffp-contract=off.
 *
 * Conventions at this file's boundary: 4x4 poses E and 12x12 Km/Dm are
 * COLUMN-MAJOR (Eigen's native layout, E[r + 4*c]); points are AoS xyz; the
 * grid is linear with index i + ni*(j + nj*k) (externals/sdfgen/array3.h:70-80).
 * Internally small matrices are row-major.
 */
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>

#define SDFO_EPS 1e-8 /* constants::eps, Utils.h:9 */

typedef struct {
    int ni, nj, nk;
    double min_box[3], max_box[3];
    double dx, dy, dz;
    double col_th;
    const double* grid;
} sdfo_grid;

/* ---------------------------------------------------------------- helpers */
static inline double g_at(const sdfo_grid* g, int i, int j, int k) {
    return g->grid[(size_t)i + (size_t)g->ni * ((size_t)j + (size_t)g->nj * (size_t)k)];
}

/* C(ar x bc) = A(ar x ac) * B(ac x bc), row-major, left-to-right accumulation */
static void mm(const double* A, int ar, int ac, const double* B, int bc, double* C) {
    for (int r = 0; r < ar; ++r)
        for (int c = 0; c < bc; ++c) {
            double s = A[r * ac] * B[c];
            for (int k = 1; k < ac; ++k) s = s + A[r * ac + k] * B[k * bc + c];
            C[r * bc + c] = s;
        }
}
static void transpose(const double* A, int r, int c, double* T) {
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) T[j * r + i] = A[i * c + j];
}
/* math::skew, Utils.h:25-33 */
static void skew3(const double v[3], double S[9]) {
    S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}
static void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static double norm3(const double a[3]) { return sqrt(dot3(a, a)); }
static double norm6(const double a[6]) {
    double s = a[0] * a[0];
    for (int i = 1; i < 6; ++i) s = s + a[i] * a[i];
    return sqrt(s);
}
/* pose pieces from a column-major 4x4 */
static void pose_Rp(const double* E, double R[9], double p[3]) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[r * 3 + c] = E[r + 4 * c];
        p[r] = E[r + 12];
    }
}
/* xw = R*xi + p : ((R_r0 x0 + R_r1 x1) + R_r2 x2) + p_r
 * CollisionDetection.cpp:136, ForceGeneralSDFContact.cpp:159,239 (order probed) */
static void to_world(const double R[9], const double p[3], const double xi[3], double xw[3]) {
    for (int r = 0; r < 3; ++r) xw[r] = ((R[r * 3] * xi[0] + R[r * 3 + 1] * xi[1]) + R[r * 3 + 2] * xi[2]) + p[r];
}
/* x = R^T (xw - p) : (R_0r d0 + R_1r d1) + R_2r d2   BodySDFObj.cpp:209-211 */
static void to_local(const double R[9], const double p[3], const double xw[3], double x[3]) {
    double d0 = xw[0] - p[0], d1 = xw[1] - p[1], d2 = xw[2] - p[2];
    for (int r = 0; r < 3; ++r) x[r] = (R[r] * d0 + R[3 + r] * d1) + R[6 + r] * d2;
}

/* ------------------------------------------------------------ SDF queries */
/* BodySDFObj::query_grid_location, BodySDFObj.cpp:379-383 */
void sdfo_grid_location(const sdfo_grid* g, const double x[3], int ijk[3]) {
    double ci = (x[0] - g->min_box[0]) / g->dx, cj = (x[1] - g->min_box[1]) / g->dy, ck = (x[2] - g->min_box[2]) / g->dz;
    ijk[0] = (int)floor(ci); ijk[1] = (int)floor(cj); ijk[2] = (int)floor(ck);
}

/* BodySDFObj::query_SDF, BodySDFObj.cpp:123-174 */
double sdfo_query_sdf(const sdfo_grid* g, const double x[3], int outside_accurate) {
    int inside = 1;
    for (int a = 0; a < 3; ++a) inside = inside && x[a] >= g->min_box[a] && x[a] <= g->max_box[a];
    if (inside) {
        double ci = (x[0] - g->min_box[0]) / g->dx, cj = (x[1] - g->min_box[1]) / g->dy, ck = (x[2] - g->min_box[2]) / g->dz;
        int i = (int)floor(ci), j = (int)floor(cj), k = (int)floor(ck);
        ci = ci - i; cj = cj - j; ck = ck - k;
        int ip = (i == g->ni - 1) ? i : i + 1;
        int jp = (j == g->nj - 1) ? j : j + 1;
        int kp = (k == g->nk - 1) ? k : k + 1;
        double v00 = (1 - ci) * g_at(g, i, j, k) + ci * g_at(g, ip, j, k);
        double v10 = (1 - ci) * g_at(g, i, jp, k) + ci * g_at(g, ip, jp, k);
        double v01 = (1 - ci) * g_at(g, i, j, kp) + ci * g_at(g, ip, j, kp);
        double v11 = (1 - ci) * g_at(g, i, jp, kp) + ci * g_at(g, ip, jp, kp);
        double w0 = (1 - cj) * v00 + cj * v10;
        double w1 = (1 - cj) * v01 + cj * v11;
        return (1 - ck) * w0 + ck * w1;
    }
    if (outside_accurate) {
        double xb[3], df[3];
        for (int a = 0; a < 3; ++a) {
            double t = x[a] > g->min_box[a] ? x[a] : g->min_box[a]; /* std::max(x, min) */
            xb[a] = t < g->max_box[a] ? t : g->max_box[a];          /* std::min(.., max) */
            df[a] = x[a] - xb[a];
        }
        return norm3(df) + sdfo_query_sdf(g, xb, 0);
    }
    return (double)FLT_MAX;
}

/* BodySDFObj::query_dSDF, BodySDFObj.cpp:176-187 */
void sdfo_query_dsdf(const sdfo_grid* g, const double x[3], double v[3]) {
    const double eps = 1e-5;
    for (int a = 0; a < 3; ++a) {
        double xp[3] = {x[0] + 0, x[1] + 0, x[2] + 0}, xm[3] = {x[0] - 0, x[1] - 0, x[2] - 0};
        xp[a] = x[a] + eps; xm[a] = x[a] - eps;
        v[a] = sdfo_query_sdf(g, xp, 1) - sdfo_query_sdf(g, xm, 1);
    }
    for (int a = 0; a < 3; ++a) v[a] = v[a] / (2 * eps);
}

/* BodySDFObj::query_ddSDF, BodySDFObj.cpp:189-205 (row-major 3x3 out) */
void sdfo_query_ddsdf(const sdfo_grid* g, const double x[3], double H[9]) {
    const double eps = 1e-5, h = eps / 2;
    double c = sdfo_query_sdf(g, x, 1);
    memset(H, 0, 9 * sizeof(double));
    for (int a = 0; a < 3; ++a) {
        double xp[3] = {x[0], x[1], x[2]}, xm[3] = {x[0], x[1], x[2]};
        xp[a] = x[a] + eps; xm[a] = x[a] - eps;
        H[a * 3 + a] = sdfo_query_sdf(g, xp, 1) + sdfo_query_sdf(g, xm, 1) - 2 * c;
    }
    static const int pr[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    for (int t = 0; t < 3; ++t) {
        int a = pr[t][0], b = pr[t][1];
        double pp[3] = {x[0], x[1], x[2]}, mm_[3] = {x[0], x[1], x[2]}, pm[3] = {x[0], x[1], x[2]}, mp[3] = {x[0], x[1], x[2]};
        pp[a] = x[a] + h; pp[b] = x[b] + h;
        mm_[a] = x[a] - h; mm_[b] = x[b] - h;
        pm[a] = x[a] + h; pm[b] = x[b] - h;
        mp[a] = x[a] - h; mp[b] = x[b] + h;
        double v = sdfo_query_sdf(g, pp, 1) + sdfo_query_sdf(g, mm_, 1) - sdfo_query_sdf(g, pm, 1) - sdfo_query_sdf(g, mp, 1);
        H[a * 3 + b] = v; H[b * 3 + a] = v;
    }
    for (int i = 0; i < 9; ++i) H[i] = H[i] / (eps * eps);
}

/* BodySDFObj::distance, BodySDFObj.cpp:207-213 */
double sdfo_distance(const sdfo_grid* g, const double* E2, const double xw[3]) {
    double R2[9], p2[3], x[3];
    pose_Rp(E2, R2, p2);
    to_local(R2, p2, xw, x);
    return sdfo_query_sdf(g, x, 0) + g->col_th;
}

/* collision_detection_general_SDF, CollisionDetection/CollisionDetection.cpp:124-147.
 * Writes per-point d, cell, world point; returns the number of active contacts
 * and their ids in point order. Any output pointer may be NULL. */
int sdfo_detect(const sdfo_grid* g, const double* E1, const double* E2, int n, const double* pts,
                double* d_out, int* cell_out, double* xw_out, int* active_ids) {
    double R1[9], p1[3], R2[9], p2[3];
    pose_Rp(E1, R1, p1); pose_Rp(E2, R2, p2);
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        double xw[3], x[3];
        to_world(R1, p1, pts + 3 * i, xw);
        to_local(R2, p2, xw, x);
        double d = sdfo_query_sdf(g, x, 0) + g->col_th;
        if (d_out) d_out[i] = d;
        if (cell_out) sdfo_grid_location(g, x, cell_out + 3 * i);
        if (xw_out) { xw_out[3 * i] = xw[0]; xw_out[3 * i + 1] = xw[1]; xw_out[3 * i + 2] = xw[2]; }
        if (d < 0.) { if (active_ids) active_ids[cnt] = i; ++cnt; }
    }
    return cnt;
}

/* ------------------------------------------------------------- collision */
typedef struct {
    double d, n[3], ddot, tdot[3], xi2[3];
    /* derivatives (row-major) */
    double dd_dxw[3], dd_dq2[6];
    double dn_dxw[9], dn_dq2[18];
    double dddot_dxw[3], dddot_dxwdot[3], dddot_dq2[6], dddot_dphi2[6];
    double dtdot_dxw[9], dtdot_dxwdot[9], dtdot_dq2[18], dtdot_dphi2[18];
    double dxi2_dxw[9], dxi2_dq2[18];
} sdfo_collision_out;

/* BodySDFObj::collision — value overload BodySDFObj.cpp:239-274, derivative
 * overload BodySDFObj.cpp:278-377 (identical values). */
void sdfo_collision(const sdfo_grid* g, const double* E2, const double phi2[6], const double xw[3],
                    const double xw_dot[3], int deriv, sdfo_collision_out* o) {
    double R2[9], p2[3], R2T[9], x[3];
    pose_Rp(E2, R2, p2);
    transpose(R2, 3, 3, R2T);
    const double* w2 = phi2; const double* v2 = phi2 + 3;
    to_local(R2, p2, xw, x);

    o->d = sdfo_query_sdf(g, x, 0) + g->col_th;
    double gr[3], e[3];
    sdfo_query_dsdf(g, x, gr);
    double gn = norm3(gr);
    for (int a = 0; a < 3; ++a) e[a] = gr[a] / gn;
    mm(R2, 3, 3, e, 1, o->n);
    for (int a = 0; a < 3; ++a) o->xi2[a] = x[a] - o->d * e[a];

    /* xdot = skew(w2)^T x + R2^T xw_dot - v2 */
    double W2[9], W2T[9], t3[3], u3[3], xdot[3];
    skew3(w2, W2); transpose(W2, 3, 3, W2T);
    mm(W2T, 3, 3, x, 1, t3);
    mm(R2T, 3, 3, xw_dot, 1, u3);
    for (int a = 0; a < 3; ++a) xdot[a] = t3[a] + u3[a] - v2[a];
    o->ddot = dot3(gr, xdot);

    double wx[3], s3[3], xw2_dot[3], vw[3];
    cross3(w2, o->xi2, wx);
    for (int a = 0; a < 3; ++a) s3[a] = wx[a] + v2[a];
    mm(R2, 3, 3, s3, 1, xw2_dot);
    for (int a = 0; a < 3; ++a) vw[a] = xw_dot[a] - xw2_dot[a];
    /* tdot = (I - n n^T) vw */
    double P[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) P[r * 3 + c] = (r == c ? 1.0 : 0.0) - o->n[r] * o->n[c];
    mm(P, 3, 3, vw, 1, o->tdot);
    if (!deriv) return;

    const double* n = o->n; const double d = o->d;
    double Sx[9];
    skew3(x, Sx);
    /* d */
    mm(gr, 1, 3, R2T, 3, o->dd_dxw);
    mm(gr, 1, 3, Sx, 3, o->dd_dq2);
    for (int a = 0; a < 3; ++a) o->dd_dq2[3 + a] = -gr[a];
    /* n */
    double de_dg[9];
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c)
            de_dg[i * 3 + c] = (gn * (i == c ? 1.0 : 0.0) - gr[i] * gr[c] / gn) / (gn * gn);
    double H[9], de_dx[9], de_dxw[9], de_dw2[9], de_dv2[9];
    sdfo_query_ddsdf(g, x, H);
    mm(de_dg, 3, 3, H, 3, de_dx);
    mm(de_dx, 3, 3, R2T, 3, de_dxw);
    mm(de_dx, 3, 3, Sx, 3, de_dw2);
    for (int i = 0; i < 9; ++i) de_dv2[i] = -de_dx[i];
    mm(R2, 3, 3, de_dxw, 3, o->dn_dxw);
    double A3[9], B3[9];
    mm(R2, 3, 3, de_dw2, 3, A3);
    for (int k = 0; k < 3; ++k) {
        double ek[3] = {0, 0, 0}, c3[3], r3[3];
        ek[k] = 1;
        cross3(ek, e, c3);         /* skew(e_k) * e */
        mm(R2, 3, 3, c3, 1, r3);
        for (int r = 0; r < 3; ++r) A3[r * 3 + k] += r3[r];
    }
    mm(R2, 3, 3, de_dv2, 3, B3);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { o->dn_dq2[r * 6 + c] = A3[r * 3 + c]; o->dn_dq2[r * 6 + 3 + c] = B3[r * 3 + c]; }
    /* xi2 */
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        o->dxi2_dxw[r * 3 + c] = R2T[r * 3 + c] - e[r] * o->dd_dxw[c] - d * de_dxw[r * 3 + c];
        o->dxi2_dq2[r * 6 + c] = Sx[r * 3 + c] - e[r] * o->dd_dq2[c] - d * de_dw2[r * 3 + c];
        o->dxi2_dq2[r * 6 + 3 + c] = -(r == c ? 1.0 : 0.0) - e[r] * o->dd_dq2[3 + c] - d * de_dv2[r * 3 + c];
    }
    /* ddot */
    double gW2T[3], xH[3], t1[3];
    mm(gr, 1, 3, W2T, 3, gW2T);
    mm(xdot, 1, 3, H, 3, xH);
    for (int a = 0; a < 3; ++a) t1[a] = gW2T[a] + xH[a];
    mm(t1, 1, 3, R2T, 3, o->dddot_dxw);
    mm(gr, 1, 3, R2T, 3, o->dddot_dxwdot);
    double WS[9], Su[9], M9[9], gM[3], xHS[3];
    mm(W2, 3, 3, Sx, 3, WS);
    skew3(u3, Su);
    for (int i = 0; i < 9; ++i) M9[i] = -WS[i] + Su[i];
    mm(gr, 1, 3, M9, 3, gM);
    mm(xH, 1, 3, Sx, 3, xHS);
    for (int a = 0; a < 3; ++a) o->dddot_dq2[a] = gM[a] + xHS[a];
    double gW2[3];
    mm(gr, 1, 3, W2, 3, gW2);
    for (int a = 0; a < 3; ++a) o->dddot_dq2[3 + a] = gW2[a] + (-xH[a]);
    mm(gr, 1, 3, Sx, 3, o->dddot_dphi2);
    for (int a = 0; a < 3; ++a) o->dddot_dphi2[3 + a] = -gr[a];
    /* xw2_dot */
    double RW[9], dxw2dot_dxw[9], dxw2dot_dq2[18], dxw2dot_dphi2[18], Ss[9], RS[9], L3[9], Rr[9], dl[9], dr[9];
    mm(R2, 3, 3, W2, 3, RW);
    mm(RW, 3, 3, o->dxi2_dxw, 3, dxw2dot_dxw);
    skew3(s3, Ss);
    mm(R2, 3, 3, Ss, 3, RS);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { dl[r * 3 + c] = o->dxi2_dq2[r * 6 + c]; dr[r * 3 + c] = o->dxi2_dq2[r * 6 + 3 + c]; }
    mm(RW, 3, 3, dl, 3, L3);
    mm(RW, 3, 3, dr, 3, Rr);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        dxw2dot_dq2[r * 6 + c] = -RS[r * 3 + c] + L3[r * 3 + c];
        dxw2dot_dq2[r * 6 + 3 + c] = Rr[r * 3 + c];
    }
    for (int k = 0; k < 3; ++k) {
        double ek[3] = {0, 0, 0}, c3[3], r3[3];
        ek[k] = 1;
        cross3(ek, o->xi2, c3);
        mm(R2, 3, 3, c3, 1, r3);
        for (int r = 0; r < 3; ++r) { dxw2dot_dphi2[r * 6 + k] = r3[r]; dxw2dot_dphi2[r * 6 + 3 + k] = R2[r * 3 + k]; }
    }
    /* tdot */
    double nvw = dot3(n, vw), vdn[3], ndx[3];
    mm(vw, 1, 3, o->dn_dxw, 3, vdn);
    mm(n, 1, 3, dxw2dot_dxw, 3, ndx);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        o->dtdot_dxw[r * 3 + c] = -dxw2dot_dxw[r * 3 + c] - nvw * o->dn_dxw[r * 3 + c] - n[r] * vdn[c] + n[r] * ndx[c];
        o->dtdot_dxwdot[r * 3 + c] = P[r * 3 + c];
    }
    double vdq[6], ndq[6];
    mm(vw, 1, 3, o->dn_dq2, 6, vdq);
    mm(n, 1, 3, dxw2dot_dq2, 6, ndq);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 6; ++c)
        o->dtdot_dq2[r * 6 + c] = -dxw2dot_dq2[r * 6 + c] - nvw * o->dn_dq2[r * 6 + c] - n[r] * (vdq[c] - ndq[c]);
    double neg[18];
    for (int i = 0; i < 18; ++i) neg[i] = -dxw2dot_dphi2[i];
    mm(P, 3, 3, neg, 6, o->dtdot_dphi2);
}

/* ------------------------------------------------------------------ force */
typedef struct {
    double kn, kt, mu, damping; /* kn, kt already multiplied by _scale * contact_scale */
} sdfo_params;

/* Gamma^T R^T for a point: 6x3 = [skew(xi); I] * R^T   (Utils.h:196-201) */
static void gtrt(const double xi[3], const double RT[9], double G[18]) {
    double S[9], top[9];
    skew3(xi, S);
    mm(S, 3, 3, RT, 3, top);
    memcpy(G, top, 9 * sizeof(double));
    memcpy(G + 9, RT, 9 * sizeof(double));
}
/* Gamma^T * skew(a): 6x3 = [skew(xi)*skew(a); skew(a)] */
static void gt_skew(const double xi[3], const double a[3], double G[18]) {
    double S[9], A[9], top[9];
    skew3(xi, S); skew3(a, A);
    mm(S, 3, 3, A, 3, top);
    memcpy(G, top, 9 * sizeof(double));
    memcpy(G + 9, A, 9 * sizeof(double));
}
#define KM(r, c) Km[(r) + 12 * (c)]
#define DM(r, c) Dm[(r) + 12 * (c)]

/* One contact of ForceGeneralSDFContact::computeForce[WithDerivative](contacts, ...)
 * value path ForceGeneralSDFContact.cpp:157-218, derivative path :237-446.
 * Accumulates into fm[12] and (deriv) column-major Km[144], Dm[144].
 * Returns 1 if the static-friction branch was taken, 0 dynamic, -1 friction off. */
static int contact_term(const sdfo_grid* g, const sdfo_params* P, const double* E1, const double phi1[6],
                        const double* E2, const double phi2[6], const double xi1[3], int deriv,
                        double* fm, double* Km, double* Dm) {
    double R1[9], p1[3], R2[9], p2[3], R1T[9], R2T[9];
    pose_Rp(E1, R1, p1); pose_Rp(E2, R2, p2);
    transpose(R1, 3, 3, R1T); transpose(R2, 3, 3, R2T);
    const double kn = P->kn, kt = P->kt, mu = P->mu, cdamp = P->damping;

    double xw1[3], S1[9], S1T[9], u[3], t3[3], xw1_dot[3];
    to_world(R1, p1, xi1, xw1);
    skew3(xi1, S1); transpose(S1, 3, 3, S1T);
    mm(S1T, 3, 3, phi1, 1, t3);
    for (int a = 0; a < 3; ++a) u[a] = t3[a] + phi1[3 + a];
    mm(R1, 3, 3, u, 1, xw1_dot);

    sdfo_collision_out c;
    sdfo_collision(g, E2, phi2, xw1, xw1_dot, deriv, &c);
    const double d = c.d, ddot = c.ddot;

    double G1[18], G2[18], ni1[6], ni2[6];
    gtrt(xi1, R1T, G1);
    gtrt(c.xi2, R2T, G2);
    mm(G1, 6, 3, c.n, 1, ni1);
    mm(G2, 6, 3, c.n, 1, ni2);

    double fc1[6], fc2[6];
    const double s1 = -kn * d + cdamp * ddot * d;
    const double s2 = kn * d - cdamp * ddot * d;
    for (int i = 0; i < 6; ++i) { fc1[i] = s1 * ni1[i]; fc2[i] = s2 * ni2[i]; fm[i] += fc1[i]; fm[6 + i] += fc2[i]; }

    double fc_norm = norm6(fc1), tdot_norm = norm3(c.tdot);
    int is_static = -1;
    if (mu > SDFO_EPS) is_static = (mu * fc_norm >= kt * tdot_norm - SDFO_EPS) ? 1 : 0;

    double G1t[6], G2t[6];
    if (mu > SDFO_EPS) {
        mm(G1, 6, 3, c.tdot, 1, G1t);
        mm(G2, 6, 3, c.tdot, 1, G2t);
    }
    if (!deriv) {
        if (mu > SDFO_EPS) {
            if (is_static) {
                for (int i = 0; i < 6; ++i) { fm[i] += -kt * G1t[i]; fm[6 + i] += kt * G2t[i]; }
            } else {
                for (int i = 0; i < 6; ++i) { fm[i] += -mu * fc_norm / tdot_norm * G1t[i]; fm[6 + i] += mu * fc_norm / tdot_norm * G2t[i]; }
            }
        }
        return is_static;
    }

    /* ---- derivatives, ForceGeneralSDFContact.cpp:242-349 ---- */
    double dxw1_dq1[18], dxw1dot_dq1[18], dxw1dot_dphi1[18], RS[9], Su[9], RSu[9], RST[9];
    mm(R1, 3, 3, S1, 3, RS);
    skew3(u, Su);
    mm(R1, 3, 3, Su, 3, RSu);
    mm(R1, 3, 3, S1T, 3, RST);
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) {
        dxw1_dq1[r * 6 + k] = -RS[r * 3 + k];      dxw1_dq1[r * 6 + 3 + k] = R1[r * 3 + k];
        dxw1dot_dq1[r * 6 + k] = -RSu[r * 3 + k];  dxw1dot_dq1[r * 6 + 3 + k] = 0;
        dxw1dot_dphi1[r * 6 + k] = RST[r * 3 + k]; dxw1dot_dphi1[r * 6 + 3 + k] = R1[r * 3 + k];
    }
    double dxi2_dq1[18];
    mm(c.dxi2_dxw, 3, 3, dxw1_dq1, 6, dxi2_dq1);

    double dd_dq1[6], dddot_dq1[6], a6[6], b6[6], dn_dq1[18], dddot_dphi1[6], dtdot_dq1[18], dtdot_dphi1[18], m18a[18], m18b[18];
    mm(c.dd_dxw, 1, 3, dxw1_dq1, 6, dd_dq1);
    mm(c.dddot_dxw, 1, 3, dxw1_dq1, 6, a6);
    mm(c.dddot_dxwdot, 1, 3, dxw1dot_dq1, 6, b6);
    for (int i = 0; i < 6; ++i) dddot_dq1[i] = a6[i] + b6[i];
    mm(c.dn_dxw, 3, 3, dxw1_dq1, 6, dn_dq1);
    mm(c.dddot_dxwdot, 1, 3, dxw1dot_dphi1, 6, dddot_dphi1);
    mm(c.dtdot_dxw, 3, 3, dxw1_dq1, 6, m18a);
    mm(c.dtdot_dxwdot, 3, 3, dxw1dot_dq1, 6, m18b);
    for (int i = 0; i < 18; ++i) dtdot_dq1[i] = m18a[i] + m18b[i];
    mm(c.dtdot_dxwdot, 3, 3, dxw1dot_dphi1, 6, dtdot_dphi1);

    double a1[3], a2[3];
    mm(R1T, 3, 3, c.n, 1, a1);
    mm(R2T, 3, 3, c.n, 1, a2);

    /* row vectors kn*dd_dq - c*dddot_dq*d - c*ddot*dd_dq */
    double rq1[6], rq2[6];
    for (int i = 0; i < 6; ++i) {
        rq1[i] = kn * dd_dq1[i] - cdamp * dddot_dq1[i] * d - cdamp * ddot * dd_dq1[i];
        rq2[i] = kn * c.dd_dq2[i] - cdamp * c.dddot_dq2[i] * d - cdamp * ddot * c.dd_dq2[i];
    }
    double G1dn1[36], G1dn2[36], G2dn1[36], G2dn2[36], G1sk[18], G2sk[18];
    mm(G1, 6, 3, dn_dq1, 6, G1dn1);
    mm(G1, 6, 3, c.dn_dq2, 6, G1dn2);
    mm(G2, 6, 3, dn_dq1, 6, G2dn1);
    mm(G2, 6, 3, c.dn_dq2, 6, G2dn2);
    gt_skew(xi1, a1, G1sk);
    gt_skew(c.xi2, a2, G2sk);

    double dfc1_dq1[36], dfc1_dq2[36], dfc1_dphi1[36], dfc1_dphi2[36];
    for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) {
        dfc1_dq1[r * 6 + k] = -ni1[r] * rq1[k] - s2 * G1dn1[r * 6 + k];
        if (k < 3) dfc1_dq1[r * 6 + k] -= s2 * G1sk[r * 3 + k];
        dfc1_dq2[r * 6 + k] = -ni1[r] * rq2[k] - s2 * G1dn2[r * 6 + k];
        dfc1_dphi1[r * 6 + k] = ni1[r] * cdamp * dddot_dphi1[k] * d;
        dfc1_dphi2[r * 6 + k] = ni1[r] * cdamp * c.dddot_dphi2[k] * d;
        KM(r, k) += dfc1_dq1[r * 6 + k];
        KM(r, 6 + k) += dfc1_dq2[r * 6 + k];
        DM(r, k) += dfc1_dphi1[r * 6 + k];
        DM(r, 6 + k) += dfc1_dphi2[r * 6 + k];
    }
    /* tmp.col(j) = dGamma2_dq(j)^T * a  =  [dxi2_dq.col(j) x a ; 0]   (.cpp:275-291,333-342) */
    double tmp1[36], tmp2[36];
    memset(tmp1, 0, sizeof tmp1); memset(tmp2, 0, sizeof tmp2);
    for (int j = 0; j < 6; ++j) {
        double c1[3] = {dxi2_dq1[j], dxi2_dq1[6 + j], dxi2_dq1[12 + j]}, c2[3] = {c.dxi2_dq2[j], c.dxi2_dq2[6 + j], c.dxi2_dq2[12 + j]}, r1[3], r2[3];
        cross3(c1, a2, r1); cross3(c2, a2, r2);
        for (int r = 0; r < 3; ++r) { tmp1[r * 6 + j] = r1[r]; tmp2[r * 6 + j] = r2[r]; }
    }
    for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) {
        KM(6 + r, k) += ni2[r] * rq1[k] + s2 * (tmp1[r * 6 + k] + G2dn1[r * 6 + k]);
        KM(6 + r, 6 + k) += ni2[r] * rq2[k] + s2 * (tmp2[r * 6 + k] + G2dn2[r * 6 + k]);
        if (k < 3) KM(6 + r, 6 + k) += s2 * G2sk[r * 3 + k];
        DM(6 + r, k) -= ni2[r] * cdamp * dddot_dphi1[k] * d;
        DM(6 + r, 6 + k) -= ni2[r] * cdamp * c.dddot_dphi2[k] * d;
    }
    if (!(mu > SDFO_EPS)) return is_static;

    /* ---- friction, ForceGeneralSDFContact.cpp:352-445 ---- */
    double dfn_dq1[6], dfn_dq2[6], dfn_dphi1[6], dfn_dphi2[6], dtn_dq1[6], dtn_dq2[6], dtn_dphi1[6], dtn_dphi2[6], r6[6];
    mm(fc1, 1, 6, dfc1_dq1, 6, r6);   for (int i = 0; i < 6; ++i) dfn_dq1[i] = (1. / fc_norm) * r6[i];
    mm(fc1, 1, 6, dfc1_dq2, 6, r6);   for (int i = 0; i < 6; ++i) dfn_dq2[i] = (1. / fc_norm) * r6[i];
    mm(fc1, 1, 6, dfc1_dphi1, 6, r6); for (int i = 0; i < 6; ++i) dfn_dphi1[i] = (1. / fc_norm) * r6[i];
    mm(fc1, 1, 6, dfc1_dphi2, 6, r6); for (int i = 0; i < 6; ++i) dfn_dphi2[i] = (1. / fc_norm) * r6[i];
    mm(c.tdot, 1, 3, dtdot_dq1, 6, r6);     for (int i = 0; i < 6; ++i) dtn_dq1[i] = (1. / tdot_norm) * r6[i];
    mm(c.tdot, 1, 3, c.dtdot_dq2, 6, r6);   for (int i = 0; i < 6; ++i) dtn_dq2[i] = (1. / tdot_norm) * r6[i];
    mm(c.tdot, 1, 3, dtdot_dphi1, 6, r6);   for (int i = 0; i < 6; ++i) dtn_dphi1[i] = (1. / tdot_norm) * r6[i];
    mm(c.tdot, 1, 3, c.dtdot_dphi2, 6, r6); for (int i = 0; i < 6; ++i) dtn_dphi2[i] = (1. / tdot_norm) * r6[i];

    double b1[3], b2[3], G1skt[18], G2skt[18];
    mm(R1T, 3, 3, c.tdot, 1, b1);
    mm(R2T, 3, 3, c.tdot, 1, b2);
    double tt1[36], tt2[36];
    memset(tt1, 0, sizeof tt1); memset(tt2, 0, sizeof tt2);
    for (int j = 0; j < 6; ++j) {
        double c1[3] = {dxi2_dq1[j], dxi2_dq1[6 + j], dxi2_dq1[12 + j]}, c2[3] = {c.dxi2_dq2[j], c.dxi2_dq2[6 + j], c.dxi2_dq2[12 + j]}, r1[3], r2[3];
        cross3(c1, b2, r1); cross3(c2, b2, r2);
        for (int r = 0; r < 3; ++r) { tt1[r * 6 + j] = r1[r]; tt2[r * 6 + j] = r2[r]; }
    }
    if (is_static) {
        double A[36], B[36], C[36], D[36], A2[36], B2[36], C2[36], D2[36];
        for (int i = 0; i < 6; ++i) { fm[i] += -kt * G1t[i]; fm[6 + i] += kt * G2t[i]; }
        mm(G1, 6, 3, dtdot_dq1, 6, A);   mm(G1, 6, 3, c.dtdot_dq2, 6, B);
        mm(G1, 6, 3, dtdot_dphi1, 6, C); mm(G1, 6, 3, c.dtdot_dphi2, 6, D);
        mm(G2, 6, 3, dtdot_dq1, 6, A2);   mm(G2, 6, 3, c.dtdot_dq2, 6, B2);
        mm(G2, 6, 3, dtdot_dphi1, 6, C2); mm(G2, 6, 3, c.dtdot_dphi2, 6, D2);
        gt_skew(xi1, b1, G1skt);
        gt_skew(c.xi2, b2, G2skt);
        for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) {
            KM(r, k) -= kt * A[r * 6 + k];
            if (k < 3) KM(r, k) -= kt * G1skt[r * 3 + k];
            KM(r, 6 + k) -= kt * B[r * 6 + k];
            DM(r, k) -= kt * C[r * 6 + k];
            DM(r, 6 + k) -= kt * D[r * 6 + k];
            KM(6 + r, k) += kt * A2[r * 6 + k] + kt * tt1[r * 6 + k];
            KM(6 + r, 6 + k) += kt * B2[r * 6 + k] + kt * tt2[r * 6 + k];
            if (k < 3) KM(6 + r, 6 + k) += kt * G2skt[r * 3 + k];
            DM(6 + r, k) += kt * C2[r * 6 + k];
            DM(6 + r, 6 + k) += kt * D2[r * 6 + k];
        }
    } else {
        const double ratio = fc_norm / tdot_norm;
        for (int i = 0; i < 6; ++i) { fm[i] += -mu * fc_norm / tdot_norm * G1t[i]; fm[6 + i] += mu * fc_norm / tdot_norm * G2t[i]; }
        /* inner(3x6) = tdot*dfn - ratio*tdot*dtn + fc_norm*dtdot */
        double in_q1[18], in_q2[18], in_p1[18], in_p2[18];
        for (int r = 0; r < 3; ++r) for (int k = 0; k < 6; ++k) {
            in_q1[r * 6 + k] = c.tdot[r] * dfn_dq1[k] - ratio * c.tdot[r] * dtn_dq1[k] + fc_norm * dtdot_dq1[r * 6 + k];
            in_q2[r * 6 + k] = c.tdot[r] * dfn_dq2[k] - ratio * c.tdot[r] * dtn_dq2[k] + fc_norm * c.dtdot_dq2[r * 6 + k];
            in_p1[r * 6 + k] = c.tdot[r] * dfn_dphi1[k] - ratio * c.tdot[r] * dtn_dphi1[k] + fc_norm * dtdot_dphi1[r * 6 + k];
            in_p2[r * 6 + k] = c.tdot[r] * dfn_dphi2[k] - ratio * c.tdot[r] * dtn_dphi2[k] + fc_norm * c.dtdot_dphi2[r * 6 + k];
        }
        double A[36], B[36], C[36], D[36], A2[36], B2[36], C2[36], D2[36];
        mm(G1, 6, 3, in_q1, 6, A); mm(G1, 6, 3, in_q2, 6, B); mm(G1, 6, 3, in_p1, 6, C); mm(G1, 6, 3, in_p2, 6, D);
        mm(G2, 6, 3, in_q1, 6, A2); mm(G2, 6, 3, in_q2, 6, B2); mm(G2, 6, 3, in_p1, 6, C2); mm(G2, 6, 3, in_p2, 6, D2);
        double sb1[3], sb2[3];
        for (int a = 0; a < 3; ++a) { sb1[a] = ratio * b1[a]; sb2[a] = ratio * b2[a]; }
        gt_skew(xi1, sb1, G1skt);
        gt_skew(c.xi2, sb2, G2skt);
        const double mt = mu / tdot_norm;
        for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) {
            KM(r, k) -= mt * A[r * 6 + k];
            if (k < 3) KM(r, k) -= mu * G1skt[r * 3 + k];
            KM(r, 6 + k) -= mt * B[r * 6 + k];
            DM(r, k) -= mt * C[r * 6 + k];
            DM(r, 6 + k) -= mt * D[r * 6 + k];
            KM(6 + r, k) += mt * A2[r * 6 + k] + mu * fc_norm / tdot_norm * tt1[r * 6 + k];
            KM(6 + r, 6 + k) += mt * B2[r * 6 + k] + mu * fc_norm / tdot_norm * tt2[r * 6 + k];
            if (k < 3) KM(6 + r, 6 + k) += mu * G2skt[r * 3 + k];
            DM(6 + r, k) += mt * C2[r * 6 + k];
            DM(6 + r, 6 + k) += mt * D2[r * 6 + k];
        }
    }
    return is_static;
}

/* One (contact body, SDF body) pair: detection + force accumulation in point
 * order, exactly the public wrappers' work minus the scatter
 * (ForceGeneralSDFContact.cpp:46-105). fm[12]; Km, Dm column-major 12x12 (may be
 * NULL when deriv == 0). counts[0] = active contacts, counts[1] = contacts that
 * took the static-friction branch. */
void sdfo_pair_force(const sdfo_grid* g, const sdfo_params* P, const double* E1, const double phi1[6],
                     const double* E2, const double phi2[6], int n, const double* pts, int deriv,
                     double* fm, double* Km, double* Dm, int* counts) {
    memset(fm, 0, 12 * sizeof(double));
    if (deriv) { memset(Km, 0, 144 * sizeof(double)); memset(Dm, 0, 144 * sizeof(double)); }
    double R1[9], p1[3], R2[9], p2[3];
    pose_Rp(E1, R1, p1); pose_Rp(E2, R2, p2);
    int n_active = 0, n_static = 0;
    for (int i = 0; i < n; ++i) {
        double xw[3], x[3];
        to_world(R1, p1, pts + 3 * i, xw);
        to_local(R2, p2, xw, x);
        double d = sdfo_query_sdf(g, x, 0) + g->col_th;
        if (!(d < 0.)) continue;
        ++n_active;
        int st = contact_term(g, P, E1, phi1, E2, phi2, pts + 3 * i, deriv, fm, Km, Dm);
        if (st == 1) ++n_static;
    }
    if (counts) { counts[0] = n_active; counts[1] = n_static; }
}

/* Per-contact terms for one point (used to validate individual contacts). */
int sdfo_contact_term(const sdfo_grid* g, const sdfo_params* P, const double* E1, const double phi1[6],
                      const double* E2, const double phi2[6], const double xi1[3], int deriv,
                      double* fm, double* Km, double* Dm) {
    memset(fm, 0, 12 * sizeof(double));
    if (deriv) { memset(Km, 0, 144 * sizeof(double)); memset(Dm, 0, 144 * sizeof(double)); }
    return contact_term(g, P, E1, phi1, E2, phi2, xi1, deriv, fm, Km, Dm);
}

/* Scatter of the pair blocks into the system arrays, ForceGeneralSDFContact.cpp:60-61,88-99.
 * fm_sys[ndof_m], Km_sys/Dm_sys column-major ndof_m x ndof_m; accumulate (+=). */
void sdfo_scatter(int ndof_m, int idx1, int idx2, const double* fm, const double* Km, const double* Dm,
                  double* fm_sys, double* Km_sys, double* Dm_sys) {
    const int idx[2] = {idx1, idx2};
    for (int b = 0; b < 2; ++b) for (int r = 0; r < 6; ++r) fm_sys[idx[b] + r] += fm[6 * b + r];
    if (!Km) return;
    for (int bc = 0; bc < 2; ++bc) for (int c = 0; c < 6; ++c)
        for (int br = 0; br < 2; ++br) for (int r = 0; r < 6; ++r) {
            size_t dst = (size_t)(idx[br] + r) + (size_t)ndof_m * (size_t)(idx[bc] + c);
            Km_sys[dst] += Km[(6 * br + r) + 12 * (6 * bc + c)];
            Dm_sys[dst] += Dm[(6 * br + r) + 12 * (6 * bc + c)];
        }
}

/* Batched helpers for the Python wrapper (avoid per-point ctypes calls). */
void sdfo_query_sdf_batch(const sdfo_grid* g, int n, const double* X, int outside_accurate, double* out) {
    for (int i = 0; i < n; ++i) out[i] = sdfo_query_sdf(g, X + 3 * i, outside_accurate);
}
void sdfo_query_dsdf_batch(const sdfo_grid* g, int n, const double* X, double* out) {
    for (int i = 0; i < n; ++i) sdfo_query_dsdf(g, X + 3 * i, out + 3 * i);
}
void sdfo_query_ddsdf_batch(const sdfo_grid* g, int n, const double* X, double* out) {
    for (int i = 0; i < n; ++i) sdfo_query_ddsdf(g, X + 3 * i, out + 9 * i);
}
void sdfo_grid_location_batch(const sdfo_grid* g, int n, const double* X, int* out) {
    for (int i = 0; i < n; ++i) sdfo_grid_location(g, X + 3 * i, out + 3 * i);
}
/* Mesh.cpp:143-198 — SDFMesh::query_SDF has no col_th and returns FLT_MAX outside;
 * SDFMesh::distance evaluates it per column. Same trilinear sample. */
void sdfo_mesh_distance_batch(const sdfo_grid* g, int n, const double* X, double* out) {
    for (int i = 0; i < n; ++i) out[i] = sdfo_query_sdf(g, X + 3 * i, 0);
}
