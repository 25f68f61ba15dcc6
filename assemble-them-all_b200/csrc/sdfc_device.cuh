// sdfc_device.cuh — device-side arithmetic of the SDF-contact hot path (sm_100a).
//
// Two groups of functions:
//   * bit-critical: pose transforms, cell index, trilinear sample and the
//     finite-difference stencil differences. They are written with the
//     round-to-nearest intrinsics (__dadd_rn / __dmul_rn / ..., which nvcc never
//     contracts into FMAs) in the evaluation order the reference's x86-64 build
//     executes, so cell indices, active masks, per-point distances and every
//     SDF sample are bit-identical to the reference:
//       Body/BodySDFObj.cpp:123-213,379-383, CollisionDetection.cpp:136-138.
//     The quotients (x - min) / dx use exact_div(): a correctly rounded division
//     from a precomputed correctly rounded reciprocal (5 FP64 instructions
//     instead of the ~20 of a generic IEEE division; same bits).
//   * contact algebra: BodySDFObj::collision (Body/BodySDFObj.cpp:239-377) and
//     the per-contact terms of ForceGeneralSDFContact::computeForce
//     [WithDerivative] (Force/ForceGeneralSDFContact.cpp:147-447), restated on
//     fixed-size register arrays. Agreement target: 1e-9 relative (fp64).
//
// Paths are relative to simulation/redmax/ of the reference.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>

namespace sdfc {

constexpr double kEps = 1e-8;     // constants::eps, Utils.h:9
constexpr double kFdEps = 1e-5;   // BodySDFObj.cpp:178,191

// Fields of BodySDFObj (Body/BodySDFObj.h:13-16) as the kernels see them.
struct GridDev {
    const double* g64;
    // fp32-lossless CELL RECORDS or nullptr: for every cell (i,j,k) the 8 corner
    // values the trilinear sample needs, contiguous and 32-byte aligned
    // (2 x float4: {000,100,010,110}, {001,101,011,111}, upper neighbours already
    // clamped as BodySDFObj.cpp:137-140 does). One query = one 32-byte sector
    // instead of four scattered row segments; 8x the footprint of the plain fp32
    // grid, which HBM capacity (180 GB) makes a non-issue.
    const float4* c32;
    // fp32-lossless BRICKED grid or nullptr (large grids, where cell records would not stay in L2):
    // the plain fp32 node values in 4 x 4 x 4 bricks of 256 contiguous bytes, brick (bx, by, bz) at
    // bx + nbx * (by + nby * bz), node (i, j, k) at (i & 3) + 4 * (j & 3) + 16 * (k & 3) inside its brick.
    // One padded node layer past n - 1 on every axis replicates the last layer, so the reference's
    // upper clamp (BodySDFObj.cpp:137-140) is a plain + 1. 4 B per node instead of 32 B per cell; a
    // detection tile stages the bricks under its bounding box in shared memory with TMA (sdfc_brick.cu),
    // every other kernel gathers the 8 corners from global memory (load_cell).
    const float* b32;
    const void* tmaps;     // device array of 16 x 16 CUtensorMap: box = 64 floats x nbx x nby x 1 bricks
    int nbx, nby, nbz;
    int ni, nj, nk;
    double mn[3], mx[3];
    double dx, dy, dz;
    double rdx, rdy, rdz;  // RN(1/dx), RN(1/dy), RN(1/dz), computed on the host with an IEEE division
    double col_th;
    // Lipschitz constant of the trilinear interpolant over the grid box: sqrt(Lx^2 + Ly^2 + Lz^2),
    // L_axis = max |difference of axis-adjacent grid values| / spacing (host, at upload). Used by
    // the stepper's conservative "skin" culling (sdfc_kernels.cu), never by the sampled values.
    double lip;
    // Smallest node value in the two outermost node layers: a lower bound of the interpolant on the
    // faces of [mn, mx] (they lie inside those layers). A point outside the box has d = FLT_MAX; to
    // become active it must first cross a face, so bmin + col_th is its margin.
    double bmin;
    __host__ __device__ bool f32() const { return c32 != nullptr || b32 != nullptr; }
};

// a / b, correctly rounded, given rb = RN(1/b). Markstein's sequence: a first
// quotient, two FMA residual corrections; the last one yields RN(a/b) because rb
// is the correctly rounded reciprocal and the corrected quotient is faithful
// (Muller et al., Handbook of Floating-Point Arithmetic, thm. on division with
// FMA). Valid without over/underflow: here a is 0 or >= 2^-60 and b is a grid
// spacing. Verified against `/` on the GPU by sdfc_selftest_exact_div().
__device__ __forceinline__ double exact_div(double a, double b, double rb) {
    const double q0 = __dmul_rn(a, rb);
    const double r0 = __fma_rn(-b, q0, a);
    const double q1 = __fma_rn(r0, rb, q0);
    const double r1 = __fma_rn(-b, q1, a);
    return __fma_rn(r1, rb, q1);
}

__device__ __forceinline__ double grid_at64(const GridDev& G, int i, int j, int k) {
    const size_t idx = (size_t)i + (size_t)G.ni * ((size_t)j + (size_t)G.nj * (size_t)k);
    return __ldg(G.g64 + idx);
}

// The 8 corners of cell (i,j,k) with the reference's upper clamp
// (BodySDFObj.cpp:137-140). Index clamps to [0,n-1] only guard memory.
// F32 = the grid is stored fp32-lossless, as cell records (c32) or as bricks (b32).
template <bool F32>
__device__ __forceinline__ void load_cell(const GridDev& G, int i, int j, int k, double v[8]) {
    i = max(0, min(i, G.ni - 1)); j = max(0, min(j, G.nj - 1)); k = max(0, min(k, G.nk - 1));
    if (F32 && G.c32 == nullptr) {   // bricked layout (uniform branch): node + 1 always exists (padded layer)
        const int i1 = i + 1, j1 = j + 1, k1 = k + 1;
        const size_t ox0 = (size_t)(i >> 2) * 64 + (i & 3), ox1 = (size_t)(i1 >> 2) * 64 + (i1 & 3);
        const size_t oy0 = (size_t)(j >> 2) * G.nbx * 64 + (j & 3) * 4, oy1 = (size_t)(j1 >> 2) * G.nbx * 64 + (j1 & 3) * 4;
        const size_t sz = (size_t)G.nbx * G.nby * 64;
        const size_t oz0 = (size_t)(k >> 2) * sz + (k & 3) * 16, oz1 = (size_t)(k1 >> 2) * sz + (k1 & 3) * 16;
        v[0] = (double)__ldg(G.b32 + ox0 + oy0 + oz0); v[1] = (double)__ldg(G.b32 + ox1 + oy0 + oz0);
        v[2] = (double)__ldg(G.b32 + ox0 + oy1 + oz0); v[3] = (double)__ldg(G.b32 + ox1 + oy1 + oz0);
        v[4] = (double)__ldg(G.b32 + ox0 + oy0 + oz1); v[5] = (double)__ldg(G.b32 + ox1 + oy0 + oz1);
        v[6] = (double)__ldg(G.b32 + ox0 + oy1 + oz1); v[7] = (double)__ldg(G.b32 + ox1 + oy1 + oz1);
        return;
    }
    if (F32) {
        const size_t idx = (size_t)i + (size_t)G.ni * ((size_t)j + (size_t)G.nj * (size_t)k);
        // one 256-bit read-only load of the 32-byte cell record (LDG.E.256 on sm_100a):
        // half the L1 wavefronts of two 128-bit gathers
#ifndef SDFC_LDG256
#define SDFC_LDG256 0  // measured on B200: the 256-bit form costs registers (occupancy) and was 20 % slower
#endif
        float4 a, b;
#if SDFC_LDG256
        asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
            : "l"(G.c32 + 2 * idx));
#else
        a = __ldg(G.c32 + 2 * idx); b = __ldg(G.c32 + 2 * idx + 1);
#endif
        v[0] = (double)a.x; v[1] = (double)a.y; v[2] = (double)a.z; v[3] = (double)a.w;
        v[4] = (double)b.x; v[5] = (double)b.y; v[6] = (double)b.z; v[7] = (double)b.w;
        return;
    }
    const int ip = (i == G.ni - 1) ? i : i + 1;
    const int jp = (j == G.nj - 1) ? j : j + 1;
    const int kp = (k == G.nk - 1) ? k : k + 1;
    v[0] = grid_at64(G, i, j, k);   v[1] = grid_at64(G, ip, j, k);
    v[2] = grid_at64(G, i, jp, k);  v[3] = grid_at64(G, ip, jp, k);
    v[4] = grid_at64(G, i, j, kp);  v[5] = grid_at64(G, ip, j, kp);
    v[6] = grid_at64(G, i, jp, kp); v[7] = grid_at64(G, ip, jp, kp);
}

// Lerp order i, then j, then k; form (1-c)*a + c*b   (BodySDFObj.cpp:143-153)
__device__ __forceinline__ double lerp1(double c, double a, double b) {
    return __dadd_rn(__dmul_rn(__dsub_rn(1.0, c), a), __dmul_rn(c, b));
}
__device__ __forceinline__ double lerp8(const double v[8], double ci, double cj, double ck) {
    const double v00 = lerp1(ci, v[0], v[1]);
    const double v10 = lerp1(ci, v[2], v[3]);
    const double v01 = lerp1(ci, v[4], v[5]);
    const double v11 = lerp1(ci, v[6], v[7]);
    const double w0 = lerp1(cj, v00, v10);
    const double w1 = lerp1(cj, v01, v11);
    return lerp1(ck, w0, w1);
}

__device__ __forceinline__ bool in_box(const GridDev& G, double x, double y, double z) {
    return x >= G.mn[0] && x <= G.mx[0] && y >= G.mn[1] && y <= G.mx[1] && z >= G.mn[2] && z <= G.mx[2];
}

// A cell whose corners are held in registers: the 7 / 26 stencil samples of an
// active contact lie within 1e-5 of x and almost always hit the same cell.
struct CellCache {
    int i, j, k;
    double v[8];
};

// query_SDF for a point known to be inside the box (BodySDFObj.cpp:130-155).
template <bool F32>
__device__ __forceinline__ double sample_inside(const GridDev& G, double x, double y, double z,
                                                const CellCache* cache) {
    double ci = exact_div(__dsub_rn(x, G.mn[0]), G.dx, G.rdx);
    double cj = exact_div(__dsub_rn(y, G.mn[1]), G.dy, G.rdy);
    double ck = exact_div(__dsub_rn(z, G.mn[2]), G.dz, G.rdz);
    const int i = (int)floor(ci), j = (int)floor(cj), k = (int)floor(ck);
    ci = __dsub_rn(ci, (double)i); cj = __dsub_rn(cj, (double)j); ck = __dsub_rn(ck, (double)k);
    if (cache != nullptr && i == cache->i && j == cache->j && k == cache->k) return lerp8(cache->v, ci, cj, ck);
    double v[8];
    load_cell<F32>(G, i, j, k, v);
    return lerp8(v, ci, cj, ck);
}

// BodySDFObj::query_SDF(x, outside_accurate)   (BodySDFObj.cpp:123-174)
template <bool F32>
__device__ __forceinline__ double query_sdf(const GridDev& G, double x, double y, double z, bool outside_accurate,
                                            const CellCache* cache) {
    if (in_box(G, x, y, z)) return sample_inside<F32>(G, x, y, z, cache);
    if (!outside_accurate) return (double)FLT_MAX;
    // closest boundary point: std::min(std::max(x, min), max)
    double bx = (x < G.mn[0]) ? G.mn[0] : x; bx = (G.mx[0] < bx) ? G.mx[0] : bx;
    double by = (y < G.mn[1]) ? G.mn[1] : y; by = (G.mx[1] < by) ? G.mx[1] : by;
    double bz = (z < G.mn[2]) ? G.mn[2] : z; bz = (G.mx[2] < bz) ? G.mx[2] : bz;
    const double ex = __dsub_rn(x, bx), ey = __dsub_rn(y, by), ez = __dsub_rn(z, bz);
    const double n2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
    return __dadd_rn(__dsqrt_rn(n2), sample_inside<F32>(G, bx, by, bz, cache));
}

// xw = R*xi + p, rows evaluated left to right   (CollisionDetection.cpp:136)
// R row-major 3x3.
__device__ __forceinline__ void to_world(const double* R, const double* p, double x0, double x1, double x2,
                                         double xw[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
        xw[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(R[r * 3], x0), __dmul_rn(R[r * 3 + 1], x1)),
                                    __dmul_rn(R[r * 3 + 2], x2)), p[r]);
}
// x = R^T (xw - p)   (BodySDFObj.cpp:209-211)
__device__ __forceinline__ void to_local(const double* R, const double* p, const double xw[3], double x[3]) {
    const double d0 = __dsub_rn(xw[0], p[0]), d1 = __dsub_rn(xw[1], p[1]), d2 = __dsub_rn(xw[2], p[2]);
#pragma unroll
    for (int r = 0; r < 3; ++r)
        x[r] = __dadd_rn(__dadd_rn(__dmul_rn(R[r], d0), __dmul_rn(R[3 + r], d1)), __dmul_rn(R[6 + r], d2));
}

// ------------------------------------------------------------ small algebra
template <int R, int K, int C>
__device__ __forceinline__ void mm(const double* A, const double* B, double* O) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            double s = A[r * K] * B[c];
#pragma unroll
            for (int k = 1; k < K; ++k) s = s + A[r * K + k] * B[k * C + c];
            O[r * C + c] = s;
        }
}
__device__ __forceinline__ void transpose3(const double* A, double* T) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T[j * 3 + i] = A[i * 3 + j];
}
__device__ __forceinline__ void skew3(const double* v, double* S) {  // math::skew, Utils.h:25-33
    S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}
__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// Poses and parameters of one (environment, pair), staged in shared memory.
struct PairPose {
    double R1[9], p1[3], R2[9], p2[3];  // row-major rotations
    double phi1[6], phi2[6];
};
struct PairParams {
    double kn, kt, mu, damping;  // kn, kt already scaled by _scale * contact_scale (.cpp:155-156)
};

// Per-contact results are handed to a Sink as soon as they are final, so that a
// warp can reduce them across its 32 contacts block by block instead of keeping
// 156-300 doubles live per thread:
//   sink.emitF(fm12)                value-only path: the 12 force entries
//   sink.emitK<B>(k36, fm3)         derivative path: 6x6 block B of Km (row-major)
//                                   plus force entries fm[3B..3B+2]
//   sink.emitD<B>(d36)              6x6 block B of Dm (only when damping/friction)
// Block order: 0 = (A,A) topLeft, 1 = (A,B) topRight, 2 = (B,A) bottomLeft, 3 = (B,B).
// Every lane of the warp must reach every emit (they contain warp shuffles).

// [skew(xi) * M ; M] for a 3xC matrix M : Gamma(xi)^T applied after R^T (Utils.h:196-201)
template <int C>
__device__ __forceinline__ void gamma_t_apply(const double* xi, const double* M, double* O /* 6xC */) {
    double S[9];
    skew3(xi, S);
    mm<3, 3, C>(S, M, O);
#pragma unroll
    for (int i = 0; i < 3 * C; ++i) O[3 * C + i] = M[i];
}

// One contact: BodySDFObj::collision + the per-contact body of
// ForceGeneralSDFContact::computeForce[WithDerivative].
//   FRIC: mu > eps (friction block .cpp:198-217 / 352-445 evaluated)
//   DAMP: damping != 0 (ddot and every Dm term are live)
// Returns 1 static friction, 0 dynamic, -1 friction off.
template <bool F32, bool DERIV, bool FRIC, bool DAMP, class Sink>
__device__ __forceinline__ int contact_eval(const GridDev& G, const PairPose& P, const PairParams& prm,
                                            const double xi1[3], Sink& sink) {
    constexpr bool DM = FRIC || DAMP;
    double fm[12];
    // With friction the blocks are revisited after the normal-force part, so all
    // of them stay live (spills); otherwise one block at a time.
    double Kst[(DERIV && FRIC) ? 4 : 1][36], Dst[(DERIV && FRIC) ? 4 : 1][36];
#define SDFC_KB(b) Kst[FRIC ? (b) : 0]
#define SDFC_DB(b) Dst[FRIC ? (b) : 0]
    const double* R1 = P.R1; const double* R2 = P.R2;
    const double* w1 = P.phi1; const double* v1 = P.phi1 + 3;
    const double* w2 = P.phi2; const double* v2 = P.phi2 + 3;
    const double kn = prm.kn, kt = prm.kt, mu = prm.mu, cd = prm.damping;
    double R1T[9], R2T[9];
    transpose3(R1, R1T); transpose3(R2, R2T);

    // ---- kinematics of the point (.cpp:158-160 / 238-240)
    double xw1[3], x[3];
    to_world(R1, P.p1, xi1[0], xi1[1], xi1[2], xw1);
    to_local(R2, P.p2, xw1, x);
    double u[3], xw1_dot[3];
    if (FRIC || DAMP) {
        double S1T[9], S1[9], t3[3];
        skew3(xi1, S1); transpose3(S1, S1T);
        mm<3, 3, 1>(S1T, w1, t3);
#pragma unroll
        for (int a = 0; a < 3; ++a) u[a] = t3[a] + v1[a];
        mm<3, 3, 1>(R1, u, xw1_dot);
    }

    // ---- SDF samples; the cell of x is cached in registers
    CellCache cc;
    {
        const double ci = exact_div(__dsub_rn(x[0], G.mn[0]), G.dx, G.rdx);
        const double cj = exact_div(__dsub_rn(x[1], G.mn[1]), G.dy, G.rdy);
        const double ck = exact_div(__dsub_rn(x[2], G.mn[2]), G.dz, G.rdz);
        cc.i = (int)floor(ci); cc.j = (int)floor(cj); cc.k = (int)floor(ck);
        load_cell<F32>(G, cc.i, cc.j, cc.k, cc.v);
    }
    const double s0 = query_sdf<F32>(G, x[0], x[1], x[2], true, &cc);  // x is inside for an active contact
    const double d = __dadd_rn(query_sdf<F32>(G, x[0], x[1], x[2], false, &cc), G.col_th);
    // query_dSDF (BodySDFObj.cpp:176-187); axis samples are reused by query_ddSDF
    double qp[3], qm[3], g[3];
    {
        qp[0] = query_sdf<F32>(G, __dadd_rn(x[0], kFdEps), x[1], x[2], true, &cc);
        qm[0] = query_sdf<F32>(G, __dsub_rn(x[0], kFdEps), x[1], x[2], true, &cc);
        qp[1] = query_sdf<F32>(G, x[0], __dadd_rn(x[1], kFdEps), x[2], true, &cc);
        qm[1] = query_sdf<F32>(G, x[0], __dsub_rn(x[1], kFdEps), x[2], true, &cc);
        qp[2] = query_sdf<F32>(G, x[0], x[1], __dadd_rn(x[2], kFdEps), true, &cc);
        qm[2] = query_sdf<F32>(G, x[0], x[1], __dsub_rn(x[2], kFdEps), true, &cc);
        constexpr double kTwoEps = 2 * kFdEps, kRTwoEps = 1.0 / kTwoEps;
#pragma unroll
        for (int a = 0; a < 3; ++a) g[a] = exact_div(__dsub_rn(qp[a], qm[a]), kTwoEps, kRTwoEps);
    }
    const double gn = sqrt(dot3(g, g));
    double e[3], n[3], xi2[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) e[a] = g[a] / gn;  // 3 IEEE divisions; feeds n, xi2 (kept exact: cheap)
    mm<3, 3, 1>(R2, e, n);
#pragma unroll
    for (int a = 0; a < 3; ++a) xi2[a] = x[a] - d * e[a];

    double xdot[3], ddot = 0, u3[3], tdot[3], vw[3], s3[3], Pn[9];
    if (FRIC || DAMP) {
        double W2T[9], W2[9], t3[3];
        skew3(w2, W2); transpose3(W2, W2T);
        mm<3, 3, 1>(W2T, x, t3);
        mm<3, 3, 1>(R2T, xw1_dot, u3);
#pragma unroll
        for (int a = 0; a < 3; ++a) xdot[a] = t3[a] + u3[a] - v2[a];
        ddot = dot3(g, xdot);
    }
    if (FRIC) {
        double wx[3], xw2_dot[3];
        cross3(w2, xi2, wx);
#pragma unroll
        for (int a = 0; a < 3; ++a) s3[a] = wx[a] + v2[a];
        mm<3, 3, 1>(R2, s3, xw2_dot);
#pragma unroll
        for (int a = 0; a < 3; ++a) vw[a] = xw1_dot[a] - xw2_dot[a];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Pn[r * 3 + c] = (r == c ? 1.0 : 0.0) - n[r] * n[c];
        mm<3, 3, 1>(Pn, vw, tdot);
    }

    // ---- contact force (.cpp:176-188 / 293-330)
    double a1[3], a2[3], ni1[6], ni2[6];
    mm<3, 3, 1>(R1T, n, a1);
    mm<3, 3, 1>(R2T, n, a2);
    gamma_t_apply<1>(xi1, a1, ni1);
    gamma_t_apply<1>(xi2, a2, ni2);
    const double s1 = DAMP ? (-kn * d + cd * ddot * d) : (-kn * d);
    const double s2 = DAMP ? (kn * d - cd * ddot * d) : (kn * d);
    double fc1[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { fc1[i] = s1 * ni1[i]; fm[i] = fc1[i]; fm[6 + i] = s2 * ni2[i]; }

    int is_static = -1;
    double fc_norm = 0, tdot_norm = 0, b1[3], b2[3], G1t[6], G2t[6];
    if (FRIC) {
        double s = fc1[0] * fc1[0];
#pragma unroll
        for (int i = 1; i < 6; ++i) s = s + fc1[i] * fc1[i];
        fc_norm = sqrt(s);
        tdot_norm = sqrt(dot3(tdot, tdot));
        is_static = (mu * fc_norm >= kt * tdot_norm - kEps) ? 1 : 0;
        mm<3, 3, 1>(R1T, tdot, b1);
        mm<3, 3, 1>(R2T, tdot, b2);
        gamma_t_apply<1>(xi1, b1, G1t);
        gamma_t_apply<1>(xi2, b2, G2t);
        if (is_static) {
#pragma unroll
            for (int i = 0; i < 6; ++i) { fm[i] += -kt * G1t[i]; fm[6 + i] += kt * G2t[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                fm[i] += -mu * fc_norm / tdot_norm * G1t[i];
                fm[6 + i] += mu * fc_norm / tdot_norm * G2t[i];
            }
        }
    }
    if constexpr (!DERIV) {
        sink.emitF(fm);
        return is_static;
    } else {
    // ================= derivatives =================
    // query_ddSDF (BodySDFObj.cpp:189-205)
    double H[9];
    {
        const double h = kFdEps / 2;
        const double two_s0 = __dmul_rn(2.0, s0);
        H[0] = __dsub_rn(__dadd_rn(qp[0], qm[0]), two_s0);
        H[4] = __dsub_rn(__dadd_rn(qp[1], qm[1]), two_s0);
        H[8] = __dsub_rn(__dadd_rn(qp[2], qm[2]), two_s0);
        const double xp = __dadd_rn(x[0], h), xm = __dsub_rn(x[0], h);
        const double yp = __dadd_rn(x[1], h), ym = __dsub_rn(x[1], h);
        const double zp = __dadd_rn(x[2], h), zm = __dsub_rn(x[2], h);
        // v = q(++) + q(--) - q(+-) - q(-+), left to right
        const double xy = __dsub_rn(__dsub_rn(__dadd_rn(query_sdf<F32>(G, xp, yp, x[2], true, &cc), query_sdf<F32>(G, xm, ym, x[2], true, &cc)),
                                              query_sdf<F32>(G, xp, ym, x[2], true, &cc)), query_sdf<F32>(G, xm, yp, x[2], true, &cc));
        const double xz = __dsub_rn(__dsub_rn(__dadd_rn(query_sdf<F32>(G, xp, x[1], zp, true, &cc), query_sdf<F32>(G, xm, x[1], zm, true, &cc)),
                                              query_sdf<F32>(G, xp, x[1], zm, true, &cc)), query_sdf<F32>(G, xm, x[1], zp, true, &cc));
        const double yz = __dsub_rn(__dsub_rn(__dadd_rn(query_sdf<F32>(G, x[0], yp, zp, true, &cc), query_sdf<F32>(G, x[0], ym, zm, true, &cc)),
                                              query_sdf<F32>(G, x[0], yp, zm, true, &cc)), query_sdf<F32>(G, x[0], ym, zp, true, &cc));
        H[1] = H[3] = xy; H[2] = H[6] = xz; H[5] = H[7] = yz;
        constexpr double kEps2 = kFdEps * kFdEps, kREps2 = 1.0 / kEps2;
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] = exact_div(H[i], kEps2, kREps2);
    }

    // ---- BodySDFObj::collision derivatives (BodySDFObj.cpp:314-376)
    double Sx[9];
    skew3(x, Sx);
    double dd_dxw[3], dd_dq2[6];
    mm<1, 3, 3>(g, R2T, dd_dxw);
    mm<1, 3, 3>(g, Sx, dd_dq2);
#pragma unroll
    for (int a = 0; a < 3; ++a) dd_dq2[3 + a] = -g[a];

    double de_dx[9], de_dxw[9], de_dw2[9];
    {
        double de_dg[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) de_dg[i * 3 + c] = (gn * (i == c ? 1.0 : 0.0) - g[i] * g[c] / gn) / (gn * gn);
        mm<3, 3, 3>(de_dg, H, de_dx);
    }
    mm<3, 3, 3>(de_dx, R2T, de_dxw);
    mm<3, 3, 3>(de_dx, Sx, de_dw2);
    double dn_dxw[9], dn_dq2[18];
    mm<3, 3, 3>(R2, de_dxw, dn_dxw);
    {
        double A3[9], B3[9], nde[9];
        mm<3, 3, 3>(R2, de_dw2, A3);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double ek[3] = {0, 0, 0}, c3[3], r3[3];
            ek[k] = 1;
            cross3(ek, e, c3);
            mm<3, 3, 1>(R2, c3, r3);
#pragma unroll
            for (int r = 0; r < 3; ++r) A3[r * 3 + k] += r3[r];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) nde[i] = -de_dx[i];
        mm<3, 3, 3>(R2, nde, B3);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) { dn_dq2[r * 6 + c] = A3[r * 3 + c]; dn_dq2[r * 6 + 3 + c] = B3[r * 3 + c]; }
    }
    double dxi2_dxw[9], dxi2_dq2[18];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            dxi2_dxw[r * 3 + c] = R2T[r * 3 + c] - e[r] * dd_dxw[c] - d * de_dxw[r * 3 + c];
            dxi2_dq2[r * 6 + c] = Sx[r * 3 + c] - e[r] * dd_dq2[c] - d * de_dw2[r * 3 + c];
            dxi2_dq2[r * 6 + 3 + c] = -(r == c ? 1.0 : 0.0) - e[r] * dd_dq2[3 + c] - d * (-de_dx[r * 3 + c]);
        }

    double dddot_dxw[3], dddot_dxwdot[3], dddot_dq2[6], dddot_dphi2[6];
    if (DAMP || FRIC) {
        double W2[9], W2T[9], gW2T[3], xH[3], t1[3];
        skew3(w2, W2); transpose3(W2, W2T);
        mm<1, 3, 3>(g, W2T, gW2T);
        mm<1, 3, 3>(xdot, H, xH);
#pragma unroll
        for (int a = 0; a < 3; ++a) t1[a] = gW2T[a] + xH[a];
        mm<1, 3, 3>(t1, R2T, dddot_dxw);
        mm<1, 3, 3>(g, R2T, dddot_dxwdot);
        double WS[9], Su[9], M9[9], gM[3], xHS[3], gW2[3];
        mm<3, 3, 3>(W2, Sx, WS);
        skew3(u3, Su);
#pragma unroll
        for (int i = 0; i < 9; ++i) M9[i] = -WS[i] + Su[i];
        mm<1, 3, 3>(g, M9, gM);
        mm<1, 3, 3>(xH, Sx, xHS);
        mm<1, 3, 3>(g, W2, gW2);
#pragma unroll
        for (int a = 0; a < 3; ++a) { dddot_dq2[a] = gM[a] + xHS[a]; dddot_dq2[3 + a] = gW2[a] + (-xH[a]); }
        mm<1, 3, 3>(g, Sx, dddot_dphi2);
#pragma unroll
        for (int a = 0; a < 3; ++a) dddot_dphi2[3 + a] = -g[a];
    }

    double dtdot_dxw[9], dtdot_dq2[18], dtdot_dphi2[18];  // dtdot_dxwdot == Pn
    if (FRIC) {
        double W2[9], RW[9], dxw2dot_dxw[9], dxw2dot_dq2[18], dxw2dot_dphi2[18];
        skew3(w2, W2);
        mm<3, 3, 3>(R2, W2, RW);
        mm<3, 3, 3>(RW, dxi2_dxw, dxw2dot_dxw);
        {
            double Ss[9], RS[9], L6[18];
            skew3(s3, Ss);
            mm<3, 3, 3>(R2, Ss, RS);
            mm<3, 3, 6>(RW, dxi2_dq2, L6);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    dxw2dot_dq2[r * 6 + c] = -RS[r * 3 + c] + L6[r * 6 + c];
                    dxw2dot_dq2[r * 6 + 3 + c] = L6[r * 6 + 3 + c];
                }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double ek[3] = {0, 0, 0}, c3[3], r3[3];
            ek[k] = 1;
            cross3(ek, xi2, c3);
            mm<3, 3, 1>(R2, c3, r3);
#pragma unroll
            for (int r = 0; r < 3; ++r) { dxw2dot_dphi2[r * 6 + k] = r3[r]; dxw2dot_dphi2[r * 6 + 3 + k] = R2[r * 3 + k]; }
        }
        const double nvw = dot3(n, vw);
        double vdn[3], ndx[3], vdq[6], ndq[6], neg[18];
        mm<1, 3, 3>(vw, dn_dxw, vdn);
        mm<1, 3, 3>(n, dxw2dot_dxw, ndx);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                dtdot_dxw[r * 3 + c] = -dxw2dot_dxw[r * 3 + c] - nvw * dn_dxw[r * 3 + c] - n[r] * vdn[c] + n[r] * ndx[c];
        mm<1, 3, 6>(vw, dn_dq2, vdq);
        mm<1, 3, 6>(n, dxw2dot_dq2, ndq);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c)
                dtdot_dq2[r * 6 + c] = -dxw2dot_dq2[r * 6 + c] - nvw * dn_dq2[r * 6 + c] - n[r] * (vdq[c] - ndq[c]);
#pragma unroll
        for (int i = 0; i < 18; ++i) neg[i] = -dxw2dot_dphi2[i];
        mm<3, 3, 6>(Pn, neg, dtdot_dphi2);
    }

    // ---- chain rule to q1 / phi1 (.cpp:242-252, 269, 305-310)
    double dxw1_dq1[18];
    {
        double S1[9], RS[9];
        skew3(xi1, S1);
        mm<3, 3, 3>(R1, S1, RS);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) { dxw1_dq1[r * 6 + k] = -RS[r * 3 + k]; dxw1_dq1[r * 6 + 3 + k] = R1[r * 3 + k]; }
    }
    double dxw1dot_dq1[18], dxw1dot_dphi1[18];
    if (DAMP || FRIC) {
        double Su[9], RSu[9], S1[9], S1T[9], RST[9];
        skew3(u, Su);
        mm<3, 3, 3>(R1, Su, RSu);
        skew3(xi1, S1); transpose3(S1, S1T);
        mm<3, 3, 3>(R1, S1T, RST);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dxw1dot_dq1[r * 6 + k] = -RSu[r * 3 + k];  dxw1dot_dq1[r * 6 + 3 + k] = 0;
                dxw1dot_dphi1[r * 6 + k] = RST[r * 3 + k]; dxw1dot_dphi1[r * 6 + 3 + k] = R1[r * 3 + k];
            }
    }
    double dxi2_dq1[18], dd_dq1[6], dn_dq1[18];
    mm<3, 3, 6>(dxi2_dxw, dxw1_dq1, dxi2_dq1);
    mm<1, 3, 6>(dd_dxw, dxw1_dq1, dd_dq1);
    mm<3, 3, 6>(dn_dxw, dxw1_dq1, dn_dq1);
    double dddot_dq1[6], dddot_dphi1[6];
    if (DAMP || FRIC) {
        double a6[6], b6[6];
        mm<1, 3, 6>(dddot_dxw, dxw1_dq1, a6);
        mm<1, 3, 6>(dddot_dxwdot, dxw1dot_dq1, b6);
#pragma unroll
        for (int i = 0; i < 6; ++i) dddot_dq1[i] = a6[i] + b6[i];
        mm<1, 3, 6>(dddot_dxwdot, dxw1dot_dphi1, dddot_dphi1);
    }
    double dtdot_dq1[18], dtdot_dphi1[18];
    if (FRIC) {
        double m18a[18], m18b[18];
        mm<3, 3, 6>(dtdot_dxw, dxw1_dq1, m18a);
        mm<3, 3, 6>(Pn, dxw1dot_dq1, m18b);
#pragma unroll
        for (int i = 0; i < 18; ++i) dtdot_dq1[i] = m18a[i] + m18b[i];
        mm<3, 3, 6>(Pn, dxw1dot_dphi1, dtdot_dphi1);
    }

    // ---- normal-force Jacobians (.cpp:312-349)
    double rq1[6], rq2[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        if (DAMP) {
            rq1[i] = kn * dd_dq1[i] - cd * dddot_dq1[i] * d - cd * ddot * dd_dq1[i];
            rq2[i] = kn * dd_dq2[i] - cd * dddot_dq2[i] * d - cd * ddot * dd_dq2[i];
        } else {
            rq1[i] = kn * dd_dq1[i];
            rq2[i] = kn * dd_dq2[i];
        }
    }
    {
        double N[18], GN[36], Sa[9], Gs[18];
        // K11 = -ni1 rq1 - s2 G1 dn_dq1 ; leftCols(3) -= s2 Gamma1^T skew(R1^T n)
        mm<3, 3, 6>(R1T, dn_dq1, N);
        gamma_t_apply<6>(xi1, N, GN);
        skew3(a1, Sa);
        gamma_t_apply<3>(xi1, Sa, Gs);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double v = -ni1[r] * rq1[k] - s2 * GN[r * 6 + k];
                if (k < 3) v -= s2 * Gs[r * 3 + k];
                SDFC_KB(0)[r * 6 + k] = v;
            }
        if constexpr (!FRIC) sink.template emitK<0>(SDFC_KB(0), fm + 0);
        // K12 = -ni1 rq2 - s2 G1 dn_dq2
        mm<3, 3, 6>(R1T, dn_dq2, N);
        gamma_t_apply<6>(xi1, N, GN);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int k = 0; k < 6; ++k) SDFC_KB(1)[r * 6 + k] = -ni1[r] * rq2[k] - s2 * GN[r * 6 + k];
        if constexpr (!FRIC) sink.template emitK<1>(SDFC_KB(1), fm + 3);
        // K21 = ni2 rq1 + s2 (tmp1 + G2 dn_dq1), tmp.col(j) = [dxi2_dq.col(j) x a2 ; 0]  (.cpp:275-291,333-338)
        mm<3, 3, 6>(R2T, dn_dq1, N);
        gamma_t_apply<6>(xi2, N, GN);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double c1[3] = {dxi2_dq1[k], dxi2_dq1[6 + k], dxi2_dq1[12 + k]}, r1[3];
            cross3(c1, a2, r1);
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const double t = (r < 3) ? r1[r % 3] : 0.0;
                SDFC_KB(2)[r * 6 + k] = ni2[r] * rq1[k] + s2 * (t + GN[r * 6 + k]);
            }
        }
        if constexpr (!FRIC) sink.template emitK<2>(SDFC_KB(2), fm + 6);
        // K22 = ni2 rq2 + s2 (tmp2 + G2 dn_dq2); leftCols(3) += s2 Gamma2^T skew(R2^T n)
        mm<3, 3, 6>(R2T, dn_dq2, N);
        gamma_t_apply<6>(xi2, N, GN);
        skew3(a2, Sa);
        gamma_t_apply<3>(xi2, Sa, Gs);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double c2[3] = {dxi2_dq2[k], dxi2_dq2[6 + k], dxi2_dq2[12 + k]}, r2[3];
            cross3(c2, a2, r2);
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const double t = (r < 3) ? r2[r % 3] : 0.0;
                double v = ni2[r] * rq2[k] + s2 * (t + GN[r * 6 + k]);
                if (k < 3) v += s2 * Gs[r * 3 + k];
                SDFC_KB(3)[r * 6 + k] = v;
            }
        }
        if constexpr (!FRIC) sink.template emitK<3>(SDFC_KB(3), fm + 9);
    }
    if constexpr (DM) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double* nn = (b < 2) ? ni1 : ni2;
            const double* dp = (b & 1) ? dddot_dphi2 : dddot_dphi1;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const double t = nn[r] * cd * dp[k] * d;
                    SDFC_DB(b)[r * 6 + k] = (b < 2) ? t : -t;
                }
            if constexpr (!FRIC) {
                if (b == 0) sink.template emitD<0>(SDFC_DB(0));
                if (b == 1) sink.template emitD<1>(SDFC_DB(1));
                if (b == 2) sink.template emitD<2>(SDFC_DB(2));
                if (b == 3) sink.template emitD<3>(SDFC_DB(3));
            }
        }
    }
    // ---- friction Jacobians (.cpp:352-445)
    if constexpr (FRIC) {
        // d|fc1|/d*, d|tdot|/d* : (1/norm) * (vec^T * Jac)
        double dfn[4][6], dtn[4][6];
        {
            double r6[6];
            mm<1, 6, 6>(fc1, Kst[0], r6);
#pragma unroll
            for (int i = 0; i < 6; ++i) dfn[0][i] = (1. / fc_norm) * r6[i];
            mm<1, 6, 6>(fc1, Kst[1], r6);
#pragma unroll
            for (int i = 0; i < 6; ++i) dfn[1][i] = (1. / fc_norm) * r6[i];
            mm<1, 6, 6>(fc1, Dst[0], r6);
#pragma unroll
            for (int i = 0; i < 6; ++i) dfn[2][i] = (1. / fc_norm) * r6[i];
            mm<1, 6, 6>(fc1, Dst[1], r6);
#pragma unroll
            for (int i = 0; i < 6; ++i) dfn[3][i] = (1. / fc_norm) * r6[i];
            mm<1, 3, 6>(tdot, dtdot_dq1, r6);
#pragma unroll
            for (int i = 0; i < 6; ++i) dtn[0][i] = (1. / tdot_norm) * r6[i];
            mm<1, 3, 6>(tdot, dtdot_dq2, r6);
#pragma unroll
            for (int i = 0; i < 6; ++i) dtn[1][i] = (1. / tdot_norm) * r6[i];
            mm<1, 3, 6>(tdot, dtdot_dphi1, r6);
#pragma unroll
            for (int i = 0; i < 6; ++i) dtn[2][i] = (1. / tdot_norm) * r6[i];
            mm<1, 3, 6>(tdot, dtdot_dphi2, r6);
#pragma unroll
            for (int i = 0; i < 6; ++i) dtn[3][i] = (1. / tdot_norm) * r6[i];
        }
        const double* J[4] = {dtdot_dq1, dtdot_dq2, dtdot_dphi1, dtdot_dphi2};
        const double ratio = fc_norm / tdot_norm;
        const double coefA = is_static ? kt : (mu / tdot_norm);         // multiplies G * inner
        const double coefT = is_static ? kt : (mu * fc_norm / tdot_norm);  // multiplies tmp (dGamma2 terms)
        const double coefS = is_static ? kt : mu;                       // multiplies Gamma^T skew(.)
#pragma unroll
        for (int v = 0; v < 4; ++v) {  // 0: dq1, 1: dq2, 2: dphi1, 3: dphi2
            double inner[18], N[18], GN1[36], GN2[36];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    if (is_static) inner[r * 6 + k] = J[v][r * 6 + k];
                    else inner[r * 6 + k] = tdot[r] * dfn[v][k] - ratio * tdot[r] * dtn[v][k] + fc_norm * J[v][r * 6 + k];
                }
            mm<3, 3, 6>(R1T, inner, N);
            gamma_t_apply<6>(xi1, N, GN1);
            mm<3, 3, 6>(R2T, inner, N);
            gamma_t_apply<6>(xi2, N, GN2);
            double* top = (v < 2) ? Kst[v] : Dst[v - 2];          // (A, .) blocks
            double* bot = (v < 2) ? Kst[2 + v] : Dst[2 + (v - 2)];  // (B, .) blocks
#pragma unroll
            for (int i = 0; i < 36; ++i) { top[i] -= coefA * GN1[i]; bot[i] += coefA * GN2[i]; }
        }
        // dGamma2 terms on the SDF body's rows (.cpp:402-412, 422-436)
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double c1[3] = {dxi2_dq1[k], dxi2_dq1[6 + k], dxi2_dq1[12 + k]}, c2[3] = {dxi2_dq2[k], dxi2_dq2[6 + k], dxi2_dq2[12 + k]};
            double r1[3], r2[3];
            cross3(c1, b2, r1); cross3(c2, b2, r2);
#pragma unroll
            for (int r = 0; r < 3; ++r) { Kst[2][r * 6 + k] += coefT * r1[r]; Kst[3][r * 6 + k] += coefT * r2[r]; }
        }
        // Gamma^T skew(R^T tdot [* ratio]) on leftCols(3) (.cpp:372,413 / 385-386,437-438)
        double sb1[3], sb2[3], Sa[9], Gs[18];
#pragma unroll
        for (int a = 0; a < 3; ++a) { sb1[a] = is_static ? b1[a] : ratio * b1[a]; sb2[a] = is_static ? b2[a] : ratio * b2[a]; }
        skew3(sb1, Sa);
        gamma_t_apply<3>(xi1, Sa, Gs);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) Kst[0][r * 6 + k] -= coefS * Gs[r * 3 + k];
        skew3(sb2, Sa);
        gamma_t_apply<3>(xi2, Sa, Gs);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) Kst[3][r * 6 + k] += coefS * Gs[r * 3 + k];
        sink.template emitK<0>(Kst[0], fm + 0);
        sink.template emitK<1>(Kst[1], fm + 3);
        sink.template emitK<2>(Kst[2], fm + 6);
        sink.template emitK<3>(Kst[3], fm + 9);
        sink.template emitD<0>(Dst[0]);
        sink.template emitD<1>(Dst[1]);
        sink.template emitD<2>(Dst[2]);
        sink.template emitD<3>(Dst[3]);
    }
    return is_static;
    }  // DERIV
#undef SDFC_KB
#undef SDFC_DB
}


// ---------------------------------------------------------------------------
// The 26 SDF samples BodySDFObj::collision takes around x (BodySDFObj.cpp:176-205,
// 254-303): centre, 6 axis samples at +-1e-5, 12 diagonal samples at +-0.5e-5 ->
// d = sdf(x) + col_th, FD gradient g, FD Hessian H, every value bit-identical to
// the reference's query_SDF sequence.
// Fast path: the stencil almost always lies inside one cell. Each axis has only 5
// distinct coordinates (0, +e, -e, +h, -h), so 15 divisions / floors / fractions
// replace 78, and the trilinear form factors stage by stage (x-lerps depend on the
// x variant only, ...): 71 lerps instead of 182 — the same operations on the same
// operands, so the same bits. Falls back to sample-by-sample evaluation when any
// coordinate leaves the centre cell or the grid box.
template <bool F32>
__device__ __noinline__ void sdf_stencil_generic(const GridDev& G, const double x[3], double& d, double g[3], double H[9]) {
    CellCache cc;
    {
        const double ci = exact_div(__dsub_rn(x[0], G.mn[0]), G.dx, G.rdx);
        const double cj = exact_div(__dsub_rn(x[1], G.mn[1]), G.dy, G.rdy);
        const double ck = exact_div(__dsub_rn(x[2], G.mn[2]), G.dz, G.rdz);
        cc.i = (int)floor(ci); cc.j = (int)floor(cj); cc.k = (int)floor(ck);
        load_cell<F32>(G, cc.i, cc.j, cc.k, cc.v);
    }
    const double s0 = query_sdf<F32>(G, x[0], x[1], x[2], true, &cc);
    d = __dadd_rn(query_sdf<F32>(G, x[0], x[1], x[2], false, &cc), G.col_th);
    double qp[3], qm[3];
    qp[0] = query_sdf<F32>(G, __dadd_rn(x[0], kFdEps), x[1], x[2], true, &cc);
    qm[0] = query_sdf<F32>(G, __dsub_rn(x[0], kFdEps), x[1], x[2], true, &cc);
    qp[1] = query_sdf<F32>(G, x[0], __dadd_rn(x[1], kFdEps), x[2], true, &cc);
    qm[1] = query_sdf<F32>(G, x[0], __dsub_rn(x[1], kFdEps), x[2], true, &cc);
    qp[2] = query_sdf<F32>(G, x[0], x[1], __dadd_rn(x[2], kFdEps), true, &cc);
    qm[2] = query_sdf<F32>(G, x[0], x[1], __dsub_rn(x[2], kFdEps), true, &cc);
    constexpr double kTwoEps = 2 * kFdEps, kRTwoEps = 1.0 / kTwoEps;
#pragma unroll
    for (int a = 0; a < 3; ++a) g[a] = exact_div(__dsub_rn(qp[a], qm[a]), kTwoEps, kRTwoEps);
    const double h = kFdEps / 2;
    const double two_s0 = __dmul_rn(2.0, s0);
    H[0] = __dsub_rn(__dadd_rn(qp[0], qm[0]), two_s0);
    H[4] = __dsub_rn(__dadd_rn(qp[1], qm[1]), two_s0);
    H[8] = __dsub_rn(__dadd_rn(qp[2], qm[2]), two_s0);
    const double xp = __dadd_rn(x[0], h), xm = __dsub_rn(x[0], h);
    const double yp = __dadd_rn(x[1], h), ym = __dsub_rn(x[1], h);
    const double zp = __dadd_rn(x[2], h), zm = __dsub_rn(x[2], h);
    const double xy = __dsub_rn(__dsub_rn(__dadd_rn(query_sdf<F32>(G, xp, yp, x[2], true, &cc), query_sdf<F32>(G, xm, ym, x[2], true, &cc)),
                                          query_sdf<F32>(G, xp, ym, x[2], true, &cc)), query_sdf<F32>(G, xm, yp, x[2], true, &cc));
    const double xz = __dsub_rn(__dsub_rn(__dadd_rn(query_sdf<F32>(G, xp, x[1], zp, true, &cc), query_sdf<F32>(G, xm, x[1], zm, true, &cc)),
                                          query_sdf<F32>(G, xp, x[1], zm, true, &cc)), query_sdf<F32>(G, xm, x[1], zp, true, &cc));
    const double yz = __dsub_rn(__dsub_rn(__dadd_rn(query_sdf<F32>(G, x[0], yp, zp, true, &cc), query_sdf<F32>(G, x[0], ym, zm, true, &cc)),
                                          query_sdf<F32>(G, x[0], yp, zm, true, &cc)), query_sdf<F32>(G, x[0], ym, zp, true, &cc));
    H[1] = H[3] = xy; H[2] = H[6] = xz; H[5] = H[7] = yz;
    constexpr double kEps2 = kFdEps * kFdEps, kREps2 = 1.0 / kEps2;
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = exact_div(H[i], kEps2, kREps2);
}

// Cell index and fractional coordinate of the 5 stencil variants of one axis:
// 0: x, 1: x+e, 2: x-e, 3: x+h, 4: x-h (BodySDFObj.cpp:132-134 per variant). base = lowest
// cell index among them, off[v] = cell(v) - base. Returns false if a variant is outside
// [mn, mx] or the variants span more than two cells.
__device__ __forceinline__ bool axis_variants(double xa, double mn, double mx, double dx, double rdx, int& base, int off[5],
                                              double f[5], int& cross) {
    constexpr double h = kFdEps / 2;
    const double t[5] = {xa, __dadd_rn(xa, kFdEps), __dsub_rn(xa, kFdEps), __dadd_rn(xa, h), __dsub_rn(xa, h)};
    bool ok = true;
    int cell[5], lo = 0x7fffffff, hi = -0x7fffffff;
#pragma unroll
    for (int v = 0; v < 5; ++v) {
        const double c = exact_div(__dsub_rn(t[v], mn), dx, rdx);
        cell[v] = (int)floor(c);
        ok = ok && (t[v] >= mn) && (t[v] <= mx);
        f[v] = __dsub_rn(c, (double)cell[v]);
        lo = min(lo, cell[v]); hi = max(hi, cell[v]);
    }
#pragma unroll
    for (int v = 0; v < 5; ++v) off[v] = cell[v] - lo;
    base = lo;
    cross = hi - lo;
    return ok && cross <= 1;
}

// The 19 distinct stencil samples on ONE cell's corner values: the trilinear form is
// evaluated stage by stage so that x-lerps are shared by all samples with the same x
// variant and so on (71 lerps instead of 19 x 7); same operations on the same
// operands as sample-by-sample evaluation, hence the same bits.
// Order: 0 centre; 1,2 x+-e; 3,4 y+-e; 5,6 z+-e; 7-10 xy (++,--,+-,-+); 11-14 xz; 15-18 yz.
__device__ __forceinline__ void stencil_on_cell(const double c[8], const double fi[5], const double fj[5], const double fk[5],
                                                double s[19]) {
    auto S = [&](int vi, int vj, int vk) {
        const double v00 = lerp1(fi[vi], c[0], c[1]);
        const double v10 = lerp1(fi[vi], c[2], c[3]);
        const double v01 = lerp1(fi[vi], c[4], c[5]);
        const double v11 = lerp1(fi[vi], c[6], c[7]);
        const double w0 = lerp1(fj[vj], v00, v10);
        const double w1 = lerp1(fj[vj], v01, v11);
        return lerp1(fk[vk], w0, w1);
    };
    s[0] = S(0, 0, 0);
    s[1] = S(1, 0, 0); s[2] = S(2, 0, 0); s[3] = S(0, 1, 0); s[4] = S(0, 2, 0); s[5] = S(0, 0, 1); s[6] = S(0, 0, 2);
    s[7] = S(3, 3, 0); s[8] = S(4, 4, 0); s[9] = S(3, 4, 0); s[10] = S(4, 3, 0);
    s[11] = S(3, 0, 3); s[12] = S(4, 0, 4); s[13] = S(3, 0, 4); s[14] = S(4, 0, 3);
    s[15] = S(0, 3, 3); s[16] = S(0, 4, 4); s[17] = S(0, 3, 4); s[18] = S(0, 4, 3);
}

// d, g, H from the 19 samples of one cell, evaluated x-variant by x-variant so that only the 8 corners, the 15
// fractions, the 4 x-lerps of the current variant and a handful of partial results are live at any time (the
// sample-ordered form above keeps all 19 samples and most stage results in registers: > 230 of them).
// Same lerps on the same operands as stencil_on_cell, so the same bits; 65 lerps.
__device__ __forceinline__ void stencil_reduce_on_cell(const double c[8], const double fi[5], const double fj[5], const double fk[5],
                                                       double col_th, double& d, double g[3], double H[9]) {
    constexpr double kTwoEps = 2 * kFdEps, kRTwoEps = 1.0 / kTwoEps;
    constexpr double kEps2 = kFdEps * kFdEps, kREps2 = 1.0 / kEps2;
    double v00, v10, v01, v11, w0, w1;
#define SDFC_ST1(vi)                                                   \
    v00 = lerp1(fi[vi], c[0], c[1]); v10 = lerp1(fi[vi], c[2], c[3]);  \
    v01 = lerp1(fi[vi], c[4], c[5]); v11 = lerp1(fi[vi], c[6], c[7])
#define SDFC_ST2(vj) w0 = lerp1(fj[vj], v00, v10); w1 = lerp1(fj[vj], v01, v11)
#define SDFC_ST3(vk) lerp1(fk[vk], w0, w1)
    // ---- x variant 0: centre, y / z axis samples, yz diagonals
    SDFC_ST1(0);
    SDFC_ST2(0);
    const double s0 = SDFC_ST3(0), s5 = SDFC_ST3(1), s6 = SDFC_ST3(2);
    SDFC_ST2(1);
    const double s3 = SDFC_ST3(0);
    SDFC_ST2(2);
    const double s4 = SDFC_ST3(0);
    SDFC_ST2(3);
    const double s15 = SDFC_ST3(3), s17 = SDFC_ST3(4);
    SDFC_ST2(4);
    const double s16 = SDFC_ST3(4), s18 = SDFC_ST3(3);
    const double two_s0 = __dmul_rn(2.0, s0);
    d = __dadd_rn(s0, col_th);
    g[1] = exact_div(__dsub_rn(s3, s4), kTwoEps, kRTwoEps);
    g[2] = exact_div(__dsub_rn(s5, s6), kTwoEps, kRTwoEps);
    H[4] = exact_div(__dsub_rn(__dadd_rn(s3, s4), two_s0), kEps2, kREps2);
    H[8] = exact_div(__dsub_rn(__dadd_rn(s5, s6), two_s0), kEps2, kREps2);
    // v = q(++) + q(--) - q(+-) - q(-+), left to right (BodySDFObj.cpp:196-203)
    H[5] = H[7] = exact_div(__dsub_rn(__dsub_rn(__dadd_rn(s15, s16), s17), s18), kEps2, kREps2);
    // ---- x variants 1, 2: x axis samples
    SDFC_ST1(1);
    SDFC_ST2(0);
    const double s1 = SDFC_ST3(0);
    SDFC_ST1(2);
    SDFC_ST2(0);
    const double s2 = SDFC_ST3(0);
    g[0] = exact_div(__dsub_rn(s1, s2), kTwoEps, kRTwoEps);
    H[0] = exact_div(__dsub_rn(__dadd_rn(s1, s2), two_s0), kEps2, kREps2);
    // ---- x variants 3, 4: xy and xz diagonals
    SDFC_ST1(3);
    SDFC_ST2(3);
    const double s7 = SDFC_ST3(0);
    SDFC_ST2(4);
    const double s9 = SDFC_ST3(0);
    SDFC_ST2(0);
    const double s11 = SDFC_ST3(3), s13 = SDFC_ST3(4);
    SDFC_ST1(4);
    SDFC_ST2(4);
    const double s8 = SDFC_ST3(0);
    SDFC_ST2(3);
    const double s10 = SDFC_ST3(0);
    SDFC_ST2(0);
    const double s12 = SDFC_ST3(4), s14 = SDFC_ST3(3);
    H[1] = H[3] = exact_div(__dsub_rn(__dsub_rn(__dadd_rn(s7, s8), s9), s10), kEps2, kREps2);
    H[2] = H[6] = exact_div(__dsub_rn(__dsub_rn(__dadd_rn(s11, s12), s13), s14), kEps2, kREps2);
#undef SDFC_ST1
#undef SDFC_ST2
#undef SDFC_ST3
}

// One-face crossing: the stencil straddles ONE cell face, so every sample lies either in the base cell (corners c)
// or in its neighbour across that face (corners c2) - decided by the sample's variant on the crossing axis alone
// (bit v of mx / my / mz: variant v of that axis is in the neighbour; at most one mask is non-zero). Same staged
// evaluation as stencil_reduce_on_cell with both cells carried through the stages that come before the crossing
// axis and a select where a stage consumes the crossing axis' variant: the x-lerps run on the cell picked by the
// x variant (x crossing) or on both cells, the y-lerps pick the x-lerps of their cell (y crossing) or run on both,
// the z-lerp picks its pair (z crossing). Each sample is the reference's lerp sequence on the reference's cell:
// same bits as 19 separate query_SDF calls (118 lerps and 140 selects instead of 133 lerps and 152 selects per
// crossing lane PLUS the 65 lerps of the other lanes of the warp, which now take this path together).
__device__ __forceinline__ void stencil_reduce_two_cells(const double c[8], const double c2[8], const double fi[5], const double fj[5],
                                                         const double fk[5], unsigned mx, unsigned my, unsigned mz, double col_th,
                                                         double& d, double g[3], double H[9]) {
    constexpr double kTwoEps = 2 * kFdEps, kRTwoEps = 1.0 / kTwoEps;
    constexpr double kEps2 = kFdEps * kFdEps, kREps2 = 1.0 / kEps2;
    double a00, a10, a01, a11, b00, b10, b01, b11, wa0, wa1, wb0, wb1;
    // x-lerps of variant vi: a* on the cell the x variant selects, b* on the neighbour (used by y / z crossings)
#define SDFC_TC1(vi)                                                                                   \
    {                                                                                                  \
        const bool nb = (mx >> (vi)) & 1u;                                                             \
        a00 = lerp1(fi[vi], nb ? c2[0] : c[0], nb ? c2[1] : c[1]);                                     \
        a10 = lerp1(fi[vi], nb ? c2[2] : c[2], nb ? c2[3] : c[3]);                                     \
        a01 = lerp1(fi[vi], nb ? c2[4] : c[4], nb ? c2[5] : c[5]);                                     \
        a11 = lerp1(fi[vi], nb ? c2[6] : c[6], nb ? c2[7] : c[7]);                                     \
        b00 = lerp1(fi[vi], c2[0], c2[1]); b10 = lerp1(fi[vi], c2[2], c2[3]);                          \
        b01 = lerp1(fi[vi], c2[4], c2[5]); b11 = lerp1(fi[vi], c2[6], c2[7]);                          \
    }
    // y-lerps of variant vj: wa* on the cell the y variant selects, wb* on the neighbour (used by z crossings)
#define SDFC_TC2(vj)                                                                                   \
    {                                                                                                  \
        const bool nb = (my >> (vj)) & 1u;                                                             \
        wa0 = lerp1(fj[vj], nb ? b00 : a00, nb ? b10 : a10);                                           \
        wa1 = lerp1(fj[vj], nb ? b01 : a01, nb ? b11 : a11);                                           \
        wb0 = lerp1(fj[vj], b00, b10); wb1 = lerp1(fj[vj], b01, b11);                                  \
    }
#define SDFC_TC3(vk) lerp1(fk[vk], ((mz >> (vk)) & 1u) ? wb0 : wa0, ((mz >> (vk)) & 1u) ? wb1 : wa1)
    // ---- x variant 0: centre, y / z axis samples, yz diagonals
    SDFC_TC1(0);
    SDFC_TC2(0);
    const double s0 = SDFC_TC3(0), s5 = SDFC_TC3(1), s6 = SDFC_TC3(2);
    SDFC_TC2(1);
    const double s3 = SDFC_TC3(0);
    SDFC_TC2(2);
    const double s4 = SDFC_TC3(0);
    SDFC_TC2(3);
    const double s15 = SDFC_TC3(3), s17 = SDFC_TC3(4);
    SDFC_TC2(4);
    const double s16 = SDFC_TC3(4), s18 = SDFC_TC3(3);
    const double two_s0 = __dmul_rn(2.0, s0);
    d = __dadd_rn(s0, col_th);
    g[1] = exact_div(__dsub_rn(s3, s4), kTwoEps, kRTwoEps);
    g[2] = exact_div(__dsub_rn(s5, s6), kTwoEps, kRTwoEps);
    H[4] = exact_div(__dsub_rn(__dadd_rn(s3, s4), two_s0), kEps2, kREps2);
    H[8] = exact_div(__dsub_rn(__dadd_rn(s5, s6), two_s0), kEps2, kREps2);
    H[5] = H[7] = exact_div(__dsub_rn(__dsub_rn(__dadd_rn(s15, s16), s17), s18), kEps2, kREps2);
    // ---- x variants 1, 2: x axis samples
    SDFC_TC1(1);
    SDFC_TC2(0);
    const double s1 = SDFC_TC3(0);
    SDFC_TC1(2);
    SDFC_TC2(0);
    const double s2 = SDFC_TC3(0);
    g[0] = exact_div(__dsub_rn(s1, s2), kTwoEps, kRTwoEps);
    H[0] = exact_div(__dsub_rn(__dadd_rn(s1, s2), two_s0), kEps2, kREps2);
    // ---- x variants 3, 4: xy and xz diagonals
    SDFC_TC1(3);
    SDFC_TC2(3);
    const double s7 = SDFC_TC3(0);
    SDFC_TC2(4);
    const double s9 = SDFC_TC3(0);
    SDFC_TC2(0);
    const double s11 = SDFC_TC3(3), s13 = SDFC_TC3(4);
    SDFC_TC1(4);
    SDFC_TC2(4);
    const double s8 = SDFC_TC3(0);
    SDFC_TC2(3);
    const double s10 = SDFC_TC3(0);
    SDFC_TC2(0);
    const double s12 = SDFC_TC3(4), s14 = SDFC_TC3(3);
    H[1] = H[3] = exact_div(__dsub_rn(__dsub_rn(__dadd_rn(s7, s8), s9), s10), kEps2, kREps2);
    H[2] = H[6] = exact_div(__dsub_rn(__dsub_rn(__dadd_rn(s11, s12), s13), s14), kEps2, kREps2);
#undef SDFC_TC1
#undef SDFC_TC2
#undef SDFC_TC3
}

// Must be called by all 32 lanes of a warp (uses a warp vote).
// Three paths, all bit-identical to the reference's 26 query_SDF calls:
//   * the stencil of every lane lies in one cell (almost always with scanned meshes): stencil_reduce_on_cell;
//   * some lane's stencil crosses one cell face along one axis (a vertex on a grid plane - common with the
//     axis-aligned synthetic parts, where whole rows of vertices sit on cell faces): the whole warp takes
//     stencil_reduce_two_cells, the lanes without a crossing with the neighbour set to the base cell;
//   * anything else (outside the box, two crossings): the sample-by-sample generic sequence.
template <bool F32>
__device__ __forceinline__ void sdf_stencil(const GridDev& G, const double x[3], double& d, double g[3], double H[9]) {
    double fi[5], fj[5], fk[5];
    int oi[5], oj[5], ok_[5], bi, bj, bk, cx, cy, cz;
    bool ok = axis_variants(x[0], G.mn[0], G.mx[0], G.dx, G.rdx, bi, oi, fi, cx);
    ok = axis_variants(x[1], G.mn[1], G.mx[1], G.dy, G.rdy, bj, oj, fj, cy) && ok;
    ok = axis_variants(x[2], G.mn[2], G.mx[2], G.dz, G.rdz, bk, ok_, fk, cz) && ok;
    ok = ok && (cx + cy + cz <= 1);
    const bool crossing = ok && (cx + cy + cz == 1);
    if (!ok) {
        double xx[3] = {x[0], x[1], x[2]}, dd, gg[3], HH[9];
        sdf_stencil_generic<F32>(G, xx, dd, gg, HH);
        d = dd;
#pragma unroll
        for (int t = 0; t < 3; ++t) g[t] = gg[t];
#pragma unroll
        for (int t = 0; t < 9; ++t) H[t] = HH[t];
    }
    if (__any_sync(0xffffffffu, crossing)) {
        if (ok) {
            unsigned mx = 0, my = 0, mz = 0;
#pragma unroll
            for (int v = 0; v < 5; ++v) { mx |= (unsigned)oi[v] << v; my |= (unsigned)oj[v] << v; mz |= (unsigned)ok_[v] << v; }
            double c[8], c2[8];
            load_cell<F32>(G, bi, bj, bk, c);
            load_cell<F32>(G, bi + cx, bj + cy, bk + cz, c2);   // the base cell again for a lane without a crossing
            stencil_reduce_two_cells(c, c2, fi, fj, fk, mx, my, mz, G.col_th, d, g, H);
        }
    } else if (ok) {
        double c[8];
        load_cell<F32>(G, bi, bj, bk, c);
        stencil_reduce_on_cell(c, fi, fj, fk, G.col_th, d, g, H);
    }
}

// ---------------------------------------------------------------------------
// Planner fast path: mu <= eps and damping == 0 (run_joint_plan.py:153), with
// derivatives. Same SDF samples as contact_eval (bit-faithful: the FD Hessian's
// round-off enters Km at ~1e-7 relative, so it must be reproduced), but the
// chain rule of BodySDFObj.cpp:314-376 + ForceGeneralSDFContact.cpp:242-349 is
// evaluated in closed form in the SDF body's frame. With
//   X   = dx/d(q1,q2) = [ -R21 [xi1] , R21 , [x] , -I ]      (3 x 12, R21 = R2^T R1)
//   P   = de/dx = (I - e e^T) H / |g| ,   a1 = R12 e ,   c = kn d
//   u1  = [xi1 x a1 ; a1] ,  u2 = [x x e ; e]
//   da1 = R12 P X + [ [a1] , 0 , -R12 [e] , 0 ]
// the reference's blocks reduce (the R2^T n = e and d-proportional terms cancel) to
//   fm  = [ -c u1 ; c u2 ]
//   K1  = -kn u1 (g X) - c [ [xi1] da1 ; da1 ]                (rows of body A)
//   K2  =  kn u2 (g X) + c [ [x] P X - [e] X ; P X ]          (rows of body B)
// Agreement with the reference's own sequence: 1e-15 relative (tests, 1e-9 bar).
// Column j of Km_sub (12 values) is final as soon as X_j is known, so it is reduced
// across the warp at once: 13 accumulators per lane instead of 156 live outputs.
struct PairPoseNF {
    double R1[9], p1[3], R2[9], p2[3];  // row-major
    double R21[9];                       // R2^T R1
};

// Constant columns factored out: writing column j of Km_sub as Y X_j + Z_j with
//   Y = [ -kn u1 g^T - c [[xi1]; I] R12 P ;  kn u2 g^T + c [ [x] P - [e] ; P ] ]     (12 x 3, per contact)
// the columns 3..5 (X_j = R21 e_m) and 9..11 (X_j = -e_m) of the SUM over contacts are (sum Y) R21 and -(sum Y):
// only the 6 pose-dependent columns 0..2, 6..8 are formed per contact, plus the three columns of Y (the general
// column with X = e_m and no explicit part) and fm. Emission (sink.emit(slot, v12)): slot 9 = fm, 6..8 = the
// columns of Y, 0..2 = Km columns 0..2, 3..5 = Km columns 6..8; the sink's owner rebuilds the 13 columns from
// the 10 summed slots (nf_slots_to_columns).
constexpr int kNfSlots = 10;

template <bool F32, class ColSink>
__device__ __forceinline__ void contact_eval_nf(const GridDev& G, const PairPoseNF& P, double kn, const double xi1[3],
                                                ColSink& sink) {
    double xw1[3], x[3];
    to_world(P.R1, P.p1, xi1[0], xi1[1], xi1[2], xw1);
    to_local(P.R2, P.p2, xw1, x);
    double d, g[3], H[9];
    sdf_stencil<F32>(G, x, d, g, H);
    const double gn = sqrt(dot3(g, g));
    const double rgn = 1.0 / gn;
    double e[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) e[a] = g[a] / gn;
    const double c = kn * d;
    // P = (H - e (e^T H)) / |g|
    double Pm[9];
    {
        double eH[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) eH[k] = e[0] * H[k] + e[1] * H[3 + k] + e[2] * H[6 + k];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) Pm[r * 3 + k] = (H[r * 3 + k] - e[r] * eH[k]) * rgn;
    }
    // a1 = R12 e = R21^T e ; u1 = [xi1 x a1 ; a1], u2 = [x x e ; e]
    double a1[3], u1[3], u2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) a1[r] = P.R21[r] * e[0] + P.R21[3 + r] * e[1] + P.R21[6 + r] * e[2];
    cross3(xi1, a1, u1); cross3(x, e, u2);
    {
        double f12[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) { f12[i] = -c * u1[i]; f12[3 + i] = -c * a1[i]; f12[6 + i] = c * u2[i]; f12[9 + i] = c * e[i]; }
        sink.emit(9, f12);
    }
    // The 9 columns in a ROLLED loop: the body (column algebra + warp reduction) is a few hundred instructions
    // that stay in the instruction cache; the unrolled form was ~70 KB of straight-line code per round and the
    // kernel stalled on instruction fetch a third of the time (ncu stall_no_inst, profiles/r2_force_kernel.md).
#pragma unroll 1
    for (int jj = 0; jj < 9; ++jj) {
        const int t = jj / 3, m = jj - 3 * t;
        const double uj[3] = {m == 0 ? 1.0 : 0.0, m == 1 ? 1.0 : 0.0, m == 2 ? 1.0 : 0.0};
        // column of X and of the explicit part of da1
        double X[3], ex[3] = {0.0, 0.0, 0.0};
        int slot;
        if (t == 0) {          // column m of Y: X = e_m
            X[0] = uj[0]; X[1] = uj[1]; X[2] = uj[2];
            slot = 6 + m;
        } else if (t == 1) {   // Km column m: X = -R21 [xi1] e_m ; extra = [a1] e_m
            double sk[3];
            cross3(xi1, uj, sk);
            cross3(a1, uj, ex);
#pragma unroll
            for (int r = 0; r < 3; ++r) X[r] = -(P.R21[r * 3] * sk[0] + P.R21[r * 3 + 1] * sk[1] + P.R21[r * 3 + 2] * sk[2]);
            slot = m;
        } else {               // Km column 6 + m: X = [x] e_m ; extra = -R12 [e] e_m
            double se[3];
            cross3(x, uj, X);
            cross3(e, uj, se);
#pragma unroll
            for (int r = 0; r < 3; ++r) ex[r] = -(P.R21[r] * se[0] + P.R21[3 + r] * se[1] + P.R21[6 + r] * se[2]);
            slot = 3 + m;
        }
        const double gX = dot3(g, X);
        double PX[3], da1[3], t1[3], t2[3], t3[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) PX[r] = Pm[r * 3] * X[0] + Pm[r * 3 + 1] * X[1] + Pm[r * 3 + 2] * X[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) da1[r] = P.R21[r] * PX[0] + P.R21[3 + r] * PX[1] + P.R21[6 + r] * PX[2] + ex[r];
        cross3(xi1, da1, t1);   // [xi1] da1
        cross3(x, PX, t2);      // [x] P X
        cross3(e, X, t3);       // [e] X
        double col[12];
        const double kg = kn * gX;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            col[r] = -kg * u1[r] - c * t1[r];
            col[3 + r] = -kg * a1[r] - c * da1[r];
            col[6 + r] = kg * u2[r] + c * (t2[r] - t3[r]);
            col[9 + r] = kg * e[r] + c * PX[r];
        }
        sink.emit(slot, col);
    }
}

// Entry `row` of the 13 columns (12 of Km_sub, then fm) from entry `row` of the 10 summed slots.
__device__ __forceinline__ void nf_slots_to_columns(const double a[kNfSlots], const double* R21, double out[13]) {
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        out[m] = a[m];
        out[6 + m] = a[3 + m];
        out[3 + m] = a[6] * R21[m] + a[7] * R21[3 + m] + a[8] * R21[6 + m];
        out[9 + m] = -a[6 + m];
    }
    out[12] = a[9];
}

}  // namespace sdfc
