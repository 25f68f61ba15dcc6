// sdfc_sdfgen.cu — SDF construction on sm_100a: a drop-in for
// sdfgen::make_level_set3 (externals/sdfgen/makelevelset3.cpp:164-226) as called by
// BodySDFObj::precompute_SDF (Body/BodySDFObj.cpp:38-84) and SDFMesh (Mesh.cpp:54-100),
// plus the grid layout rule and the `.sdf` text cache (BodySDFObj.cpp:86-116).
//
// The reference's result depends on its sequential order in two places; both are kept
// exactly, so the grid is BIT-IDENTICAL to the CPU one (contact parity depends on it):
//   * exact band (.cpp:175-192): triangles are visited in order and a cell keeps the first
//     triangle that reaches its minimum distance (strict <). Here every (triangle, cell)
//     pair is evaluated in parallel and combined with a 64-bit atomicMin on
//     (float bits of d, triangle index): the minimum distance, lowest index among ties.
//   * fast sweeping (.cpp:99-124,212-222): 2 x 8 Gauss-Seidel sweeps, each cell reading
//     the closest triangles of its 7 upwind neighbours already updated in the same sweep.
//     Upwind neighbours lie on earlier anti-diagonal planes i'+j'+k' = s of the sweep's own
//     index order, so processing plane after plane reproduces the sequential result; the planes
//     are separated by grid-wide barriers inside one persistent cooperative kernel.
//   * parity signs (.cpp:193-208,223-225): integer intersection counts (atomicAdd) and a
//     prefix parity along x.
// Arithmetic: fp32 distances, fp64 grid coordinates / barycentrics, in the reference's
// evaluation order; this translation unit is compiled with -fmad=false (the reference's
// x86-64 build has no FMA) and nvcc's default IEEE division and square root.
#include "sdfc_host.h"

#include <cooperative_groups.h>

#include <cfloat>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <string>

namespace sdfc {

struct F3 { float v[3]; };

struct LevelSetDev {
    const unsigned* tri;  // [n_tri][3]
    const float* x;       // [n_vert][3]
    int n_tri;
    float o[3], dx, dy, dz;
    int ni, nj, nk, band;
    unsigned long long* key;  // per cell: (bits(d) << 32) | triangle
    int* count;               // intersection_count
    float* phi;
    int* closest;
    unsigned init_bits;
};

__device__ __forceinline__ F3 ld3(const float* x, unsigned i) { F3 r = {{x[3 * i], x[3 * i + 1], x[3 * i + 2]}}; return r; }
__device__ __forceinline__ F3 sub3(F3 a, F3 b) { F3 r = {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; return r; }
__device__ __forceinline__ float dot3(F3 a, F3 b) { float d = a.v[0] * b.v[0]; d += a.v[1] * b.v[1]; d += a.v[2] * b.v[2]; return d; }
__device__ __forceinline__ float mag2(F3 a) { float l = a.v[0] * a.v[0]; l += a.v[1] * a.v[1]; l += a.v[2] * a.v[2]; return l; }
__device__ __forceinline__ float dist3(F3 a, F3 b) {  // vec.h:226-235
    float t = a.v[0] - b.v[0];
    float d = t * t;
    t = a.v[1] - b.v[1]; d += t * t;
    t = a.v[2] - b.v[2]; d += t * t;
    return sqrtf(d);
}
// makelevelset3.cpp:5-24
__device__ __forceinline__ float seg_dist(F3 x0, F3 x1, F3 x2) {
    const F3 dx = sub3(x2, x1);
    const double m2 = (double)mag2(dx);
    float s12 = (float)((double)dot3(sub3(x2, x0), dx) / m2);
    if (s12 < 0) s12 = 0;
    else if (s12 > 1) s12 = 1;
    const float o = 1 - s12;
    const F3 c = {{s12 * x1.v[0] + o * x2.v[0], s12 * x1.v[1] + o * x2.v[1], s12 * x1.v[2] + o * x2.v[2]}};
    return dist3(x0, c);
}
// makelevelset3.cpp:26-48
__device__ float tri_dist(F3 x0, F3 x1, F3 x2, F3 x3) {
    const F3 x13 = sub3(x1, x3), x23 = sub3(x2, x3), x03 = sub3(x0, x3);
    const float m13 = mag2(x13), m23 = mag2(x23), d = dot3(x13, x23);
    const float det = m13 * m23 - d * d;
    const float invdet = 1.f / (det < 1e-30f ? 1e-30f : det);
    const float a = dot3(x13, x03), b = dot3(x23, x03);
    const float w23 = invdet * (m23 * a - d * b);
    const float w31 = invdet * (m13 * b - d * a);
    const float w12 = 1 - w23 - w31;
    if (w23 >= 0 && w31 >= 0 && w12 >= 0) {
        const F3 c = {{w23 * x1.v[0] + w31 * x2.v[0] + w12 * x3.v[0], w23 * x1.v[1] + w31 * x2.v[1] + w12 * x3.v[1],
                       w23 * x1.v[2] + w31 * x2.v[2] + w12 * x3.v[2]}};
        return dist3(x0, c);
    }
    float p, q;
    if (w23 > 0) { p = seg_dist(x0, x1, x2); q = seg_dist(x0, x1, x3); }
    else if (w31 > 0) { p = seg_dist(x0, x1, x2); q = seg_dist(x0, x2, x3); }
    else { p = seg_dist(x0, x1, x3); q = seg_dist(x0, x2, x3); }
    return (q < p) ? q : p;
}

// makelevelset3.cpp:126-162
__device__ __forceinline__ int orientation(double x1, double y1, double x2, double y2, double& area) {
    area = y1 * x2 - x1 * y2;
    if (area > 0) return 1;
    else if (area < 0) return -1;
    else if (y2 > y1) return 1;
    else if (y2 < y1) return -1;
    else if (x1 > x2) return 1;
    else if (x1 < x2) return -1;
    else return 0;
}
__device__ __forceinline__ bool in_tri_2d(double x0, double y0, double x1, double y1, double x2, double y2, double x3,
                                          double y3, double& a, double& b, double& c) {
    x1 -= x0; x2 -= x0; x3 -= x0;
    y1 -= y0; y2 -= y0; y3 -= y0;
    const int sa = orientation(x2, y2, x3, y3, a);
    if (sa == 0) return false;
    const int sb = orientation(x3, y3, x1, y1, b);
    if (sb != sa) return false;
    const int sc = orientation(x1, y1, x2, y2, c);
    if (sc != sa) return false;
    const double sum = a + b + c;
    a /= sum; b /= sum; c /= sum;
    return true;
}

__device__ __forceinline__ double min3(double a, double b, double c) { const double m = a < b ? a : b; return m < c ? m : c; }
__device__ __forceinline__ double max3(double a, double b, double c) { const double m = a > b ? a : b; return m > c ? m : c; }
__device__ __forceinline__ int clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }
__device__ __forceinline__ size_t cell(const LevelSetDev& L, int i, int j, int k) {
    return (size_t)i + (size_t)L.ni * ((size_t)j + (size_t)L.nj * (size_t)k);
}
__device__ __forceinline__ F3 grid_point(const LevelSetDev& L, int i, int j, int k) {
    F3 g = {{i * L.dx + L.o[0], j * L.dy + L.o[1], k * L.dz + L.o[2]}};
    return g;
}

__global__ void k_ls_init(LevelSetDev L, size_t n) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    L.key[c] = ((unsigned long long)L.init_bits << 32) | 0xFFFFFFFFull;
    L.count[c] = 0;
}

// One CTA per triangle: exact band distances and x-ray intersection counts.
__global__ void __launch_bounds__(128) k_ls_band(LevelSetDev L) {
    const int t = blockIdx.x;
    const unsigned p = L.tri[3 * t], q = L.tri[3 * t + 1], r = L.tri[3 * t + 2];
    const F3 xp = ld3(L.x, p), xq = ld3(L.x, q), xr = ld3(L.x, r);
    const double fip = ((double)xp.v[0] - L.o[0]) / L.dx, fjp = ((double)xp.v[1] - L.o[1]) / L.dy, fkp = ((double)xp.v[2] - L.o[2]) / L.dz;
    const double fiq = ((double)xq.v[0] - L.o[0]) / L.dx, fjq = ((double)xq.v[1] - L.o[1]) / L.dy, fkq = ((double)xq.v[2] - L.o[2]) / L.dz;
    const double fir = ((double)xr.v[0] - L.o[0]) / L.dx, fjr = ((double)xr.v[1] - L.o[1]) / L.dy, fkr = ((double)xr.v[2] - L.o[2]) / L.dz;
    {
        const int i0 = clampi((int)min3(fip, fiq, fir) - L.band, 0, L.ni - 1), i1 = clampi((int)max3(fip, fiq, fir) + L.band + 1, 0, L.ni - 1);
        const int j0 = clampi((int)min3(fjp, fjq, fjr) - L.band, 0, L.nj - 1), j1 = clampi((int)max3(fjp, fjq, fjr) + L.band + 1, 0, L.nj - 1);
        const int k0 = clampi((int)min3(fkp, fkq, fkr) - L.band, 0, L.nk - 1), k1 = clampi((int)max3(fkp, fkq, fkr) + L.band + 1, 0, L.nk - 1);
        const int wi = i1 - i0 + 1, wj = j1 - j0 + 1, wk = k1 - k0 + 1;
        const long long total = (long long)wi * wj * wk;
        for (long long c = threadIdx.x; c < total; c += blockDim.x) {
            const int i = i0 + (int)(c % wi), j = j0 + (int)((c / wi) % wj), k = k0 + (int)(c / ((long long)wi * wj));
            const float d = tri_dist(grid_point(L, i, j, k), xp, xq, xr);
            // d >= 0: the bit pattern orders like the value; ties keep the lowest triangle
            atomicMin(L.key + cell(L, i, j, k), ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)t);
        }
    }
    {
        const int j0 = clampi((int)ceil(min3(fjp, fjq, fjr)), 0, L.nj - 1), j1 = clampi((int)floor(max3(fjp, fjq, fjr)), 0, L.nj - 1);
        const int k0 = clampi((int)ceil(min3(fkp, fkq, fkr)), 0, L.nk - 1), k1 = clampi((int)floor(max3(fkp, fkq, fkr)), 0, L.nk - 1);
        const int wj = j1 - j0 + 1, wk = k1 - k0 + 1;
        if (wj <= 0 || wk <= 0) return;
        for (int c = threadIdx.x; c < wj * wk; c += blockDim.x) {
            const int j = j0 + c % wj, k = k0 + c / wj;
            double a, b, cc;
            if (in_tri_2d(j, k, fjp, fkp, fjq, fkq, fjr, fkr, a, b, cc)) {
                const double fi = a * fip + b * fiq + cc * fir;
                const int ii = (int)ceil(fi);
                if (ii < 0) atomicAdd(L.count + cell(L, 0, j, k), 1);
                else if (ii < L.ni) atomicAdd(L.count + cell(L, ii, j, k), 1);
            }
        }
    }
}

__global__ void k_ls_decode(LevelSetDev L, size_t n) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const unsigned long long key = L.key[c];
    const unsigned hi = (unsigned)(key >> 32);
    // a triangle at exactly the initial bound never passes the reference's strict `d < phi`
    const bool hit = hi != L.init_bits;
    L.phi[c] = __uint_as_float(hi);
    L.closest[c] = hit ? (int)(unsigned)(key & 0xFFFFFFFFull) : -1;
}

// One cell of one sweep (makelevelset3.cpp:86-97,112-123). Primed indices count along the
// sweep direction starting at 0; the 7 upwind neighbours lie on earlier planes i'+j'+k'.
__device__ __forceinline__ void sweep_cell(const LevelSetDev& L, int ip, int jp, int kp, int di, int dj, int dk) {
    const int i = di > 0 ? ip : L.ni - 1 - ip, j = dj > 0 ? jp : L.nj - 1 - jp, k = dk > 0 ? kp : L.nk - 1 - kp;
    const size_t c0 = cell(L, i, j, k);
    // phi / closest are rewritten plane after plane by other CTAs of the same (persistent) kernel:
    // read them through L2 (ld.global.cg), never from a possibly stale L1 line
    float best = __ldcg(L.phi + c0);
    int best_t = __ldcg(L.closest + c0);
    const F3 gx = grid_point(L, i, j, k);
    const int ni[7] = {i - di, i, i - di, i, i - di, i, i - di};
    const int nj[7] = {j, j - dj, j - dj, j, j, j - dj, j - dj};
    const int nk[7] = {k, k, k, k - dk, k - dk, k - dk, k - dk};
    int last_t = -1;
    float last_d = 0.f;
#pragma unroll
    for (int n = 0; n < 7; ++n) {
        const int t = __ldcg(L.closest + cell(L, ni[n], nj[n], nk[n]));
        if (t < 0) continue;
        float d;
        if (t == last_t) d = last_d;  // same triangle, same arithmetic: identical value
        else {
            d = tri_dist(gx, ld3(L.x, L.tri[3 * t]), ld3(L.x, L.tri[3 * t + 1]), ld3(L.x, L.tri[3 * t + 2]));
            last_t = t; last_d = d;
        }
        if (d < best) { best = d; best_t = t; }
    }
    L.phi[c0] = best;
    L.closest[c0] = best_t;
}

// One anti-diagonal plane of one sweep: threads over (j', k'), i' = s - j' - k'. Fallback path
// (one launch per plane) when a cooperative launch is not available.
__global__ void __launch_bounds__(128) k_ls_sweep(LevelSetDev L, int s, int di, int dj, int dk) {
    const int jp = 1 + blockIdx.x * blockDim.x + threadIdx.x, kp = 1 + blockIdx.y;
    if (jp > L.nj - 1 || kp > L.nk - 1) return;
    const int ip = s - jp - kp;
    if (ip < 1 || ip > L.ni - 1) return;
    sweep_cell(L, ip, jp, kp, di, dj, dk);
}

// All 2 x 8 sweeps in ONE persistent cooperative kernel: the planes of a sweep are separated by
// grid-wide barriers instead of kernel boundaries (~4.7 k planes for a 100^3 grid).
__global__ void __launch_bounds__(256) k_ls_sweep_all(LevelSetDev L) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int dirs[8][3] = {{1, 1, 1}, {-1, -1, -1}, {1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, -1}, {1, -1, -1}, {-1, 1, 1}};
    const int wj = L.nj - 1, wk = L.nk - 1;
    const int s_max = (L.ni - 1) + wj + wk;
    const long long gsize = (long long)gridDim.x * blockDim.x, gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int pass = 0; pass < 2; ++pass)
        for (int d = 0; d < 8; ++d)
            for (int s = 3; s <= s_max; ++s) {
                // (j', k') pairs of this plane: k' in [max(1, s - (ni-1) - wj), min(wk, s - 2)]
                const int k_lo = max(1, s - (L.ni - 1) - wj), k_hi = min(wk, s - 2);
                const long long span = (long long)(k_hi - k_lo + 1) * wj;
                for (long long idx = gtid; idx < span; idx += gsize) {
                    const int jp = 1 + (int)(idx % wj), kp = k_lo + (int)(idx / wj);
                    const int ip = s - jp - kp;
                    if (ip >= 1 && ip <= L.ni - 1) sweep_cell(L, ip, jp, kp, dirs[d][0], dirs[d][1], dirs[d][2]);
                }
                grid.sync();
            }
}

__global__ void k_ls_sign(LevelSetDev L) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (j >= L.nj || k >= L.nk) return;
    int total = 0;
    for (int i = 0; i < L.ni; ++i) {
        const size_t c = cell(L, i, j, k);
        total += L.count[c];
        if (total % 2 == 1) L.phi[c] = -L.phi[c];
    }
}

}  // namespace sdfc

using namespace sdfc;
using sdfc_host::dev_alloc;
using sdfc_host::fail;

extern "C" {

sdfc_status sdfc_make_level_set3(sdfc_context* ctx, int n_tri, const uint32_t* tri, int n_vert, const float* x,
                                 const float origin[3], float dx, float dy, float dz, int ni, int nj, int nk,
                                 int exact_band, float* phi_out) {
    SDFC_REQUIRE(ctx && tri && x && origin && phi_out, "sdfc_make_level_set3: NULL argument");
    SDFC_REQUIRE(n_tri > 0 && n_vert > 0 && ni > 1 && nj > 1 && nk > 1 && exact_band >= 0, "sdfc_make_level_set3: bad size");
    for (int i = 0; i < 3 * n_tri; ++i) SDFC_REQUIRE(tri[i] < (uint32_t)n_vert, "sdfc_make_level_set3: vertex index out of range");
    SDFC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)ni * nj * nk;
    LevelSetDev L{};
    unsigned* d_tri = nullptr; float* d_x = nullptr;
    sdfc_status rc = SDFC_OK;
    auto A = [&](sdfc_status s) { if (rc == SDFC_OK) rc = s; };
    A(dev_alloc(&d_tri, (size_t)3 * n_tri));
    A(dev_alloc(&d_x, (size_t)3 * n_vert));
    A(dev_alloc(&L.key, n));
    A(dev_alloc(&L.count, n));
    A(dev_alloc(&L.phi, n));
    A(dev_alloc(&L.closest, n));
    auto cleanup = [&]() { cudaFree(d_tri); cudaFree(d_x); cudaFree(L.key); cudaFree(L.count); cudaFree(L.phi); cudaFree(L.closest); };
    if (rc != SDFC_OK) { cleanup(); return rc; }
    L.tri = d_tri; L.x = d_x; L.n_tri = n_tri;
    L.o[0] = origin[0]; L.o[1] = origin[1]; L.o[2] = origin[2];
    L.dx = dx; L.dy = dy; L.dz = dz; L.ni = ni; L.nj = nj; L.nk = nk; L.band = exact_band;
    const float init = (ni + nj + nk) * (dx + dy + dz);  // makelevelset3.cpp:169
    std::memcpy(&L.init_bits, &init, sizeof(float));
    cudaError_t e = cudaMemcpyAsync(d_tri, tri, sizeof(unsigned) * 3 * n_tri, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, x, sizeof(float) * 3 * n_vert, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        const unsigned nb = (unsigned)((n + 255) / 256);
        k_ls_init<<<nb, 256, 0, st>>>(L, n);
        k_ls_band<<<n_tri, 128, 0, st>>>(L);
        k_ls_decode<<<nb, 256, 0, st>>>(L, n);
        ctx->launches += 3;
        // sweeps: one persistent cooperative kernel (grid-wide barrier per plane) when the device
        // supports it, else one launch per plane
        int coop = 0, per_sm = 0, sms = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        if (coop) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ls_sweep_all, 256, 0);
        bool done = false;
        if (coop && per_sm > 0 && std::getenv("SDFC_SDFGEN_NO_COOP") == nullptr) {
            // enough CTAs for the widest plane, never more than can be co-resident
            const long long widest = (long long)(nj - 1) * (nk - 1);
            const int want = (int)std::min<long long>((widest + 255) / 256, (long long)sms * std::min(per_sm, 2));
            void* args[] = {(void*)&L};
            if (cudaLaunchCooperativeKernel((void*)k_ls_sweep_all, dim3(std::max(want, 1)), dim3(256), args, 0, st) == cudaSuccess) {
                done = true;
                ctx->launches += 1;
            } else {
                cudaGetLastError();
            }
        }
        if (!done) {
            const dim3 sg((nj - 1 + 127) / 128, nk - 1);
            static const int dirs[8][3] = {{1, 1, 1}, {-1, -1, -1}, {1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, -1}, {1, -1, -1}, {-1, 1, 1}};
            const int s_max = (ni - 1) + (nj - 1) + (nk - 1);
            for (int pass = 0; pass < 2; ++pass)
                for (int d = 0; d < 8; ++d)
                    for (int s = 3; s <= s_max; ++s) {
                        k_ls_sweep<<<sg, 128, 0, st>>>(L, s, dirs[d][0], dirs[d][1], dirs[d][2]);
                        ctx->launches += 1;
                    }
        }
        k_ls_sign<<<dim3((nj + 127) / 128, nk), 128, 0, st>>>(L);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(phi_out, L.phi, sizeof(float) * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess) return fail(SDFC_ERR_CUDA, std::string("sdfc_make_level_set3: ") + cudaGetErrorString(e));
    return SDFC_OK;
}

// BodySDFObj::precompute_SDF, lines 41-70: spacing, padded fp32 box, grid size.
sdfc_status sdfc_sdf_layout(int n_vert, const double* V, double dx, int res, double spacing[3], float min_box[3],
                            float max_box[3], int32_t dims[3]) {
    SDFC_REQUIRE(V && spacing && min_box && max_box && dims, "sdfc_sdf_layout: NULL argument");
    SDFC_REQUIRE(n_vert > 0 && dx > 0 && res > 0, "sdfc_sdf_layout: bad argument");
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        min_box[a] = std::numeric_limits<float>::max();
        max_box[a] = -std::numeric_limits<float>::max();
        lo[a] = hi[a] = V[a];
    }
    for (int i = 0; i < n_vert; ++i)
        for (int a = 0; a < 3; ++a) {
            const double v = V[3 * (size_t)i + a];
            const float p = (float)v;
            if (p < min_box[a]) min_box[a] = p;
            if (p > max_box[a]) max_box[a] = p;
            if (v < lo[a]) lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    for (int a = 0; a < 3; ++a) {
        const double ext = (hi[a] - lo[a]) / res;
        spacing[a] = std::min(dx, ext);
        const float unit = (float)spacing[a];
        min_box[a] -= 1.f * unit;
        max_box[a] += 1.f * unit;
        const float half = (float)(0.5 * spacing[a]);
        const float num = max_box[a] - min_box[a] + half;
        dims[a] = (int32_t)(unsigned)(num / (float)spacing[a]) + 1;
    }
    return SDFC_OK;
}

// BodySDFObj::save_SDF / load_SDF (BodySDFObj.cpp:86-116): the `.sdf` text cache, written
// with the stream's default 6 significant digits exactly like the reference.
sdfc_status sdfc_sdf_save_text(const char* path, int ni, int nj, int nk, const double min_box[3], const double max_box[3],
                               double dx, double dy, double dz, const double* grid) {
    SDFC_REQUIRE(path && min_box && max_box && grid, "sdfc_sdf_save_text: NULL argument");
    std::ofstream out(path);
    if (!out.good()) return fail(SDFC_ERR_INVALID, std::string("sdfc_sdf_save_text: cannot open ") + path);
    out << ni << " " << nj << " " << nk << std::endl;
    out << min_box[0] << " " << min_box[1] << " " << min_box[2] << std::endl;
    out << max_box[0] << " " << max_box[1] << " " << max_box[2] << std::endl;
    out << dx << " " << dy << " " << dz << std::endl;
    const size_t n = (size_t)ni * nj * nk;
    for (size_t i = 0; i < n; ++i) out << grid[i] << std::endl;
    out.close();
    return SDFC_OK;
}

sdfc_status sdfc_sdf_load_text_header(const char* path, int32_t dims[3]) {
    SDFC_REQUIRE(path && dims, "sdfc_sdf_load_text_header: NULL argument");
    std::ifstream in(path);
    if (!in.good()) return fail(SDFC_ERR_INVALID, std::string("sdfc_sdf_load_text_header: cannot open ") + path);
    in >> dims[0] >> dims[1] >> dims[2];
    return in.fail() ? fail(SDFC_ERR_INVALID, "sdfc_sdf_load_text_header: malformed header") : SDFC_OK;
}

sdfc_status sdfc_sdf_load_text(const char* path, double min_box[3], double max_box[3], double spacing[3], double* grid,
                               int64_t capacity) {
    SDFC_REQUIRE(path && min_box && max_box && spacing && grid, "sdfc_sdf_load_text: NULL argument");
    std::ifstream in(path);
    if (!in.good()) return fail(SDFC_ERR_INVALID, std::string("sdfc_sdf_load_text: cannot open ") + path);
    int ni, nj, nk;
    in >> ni >> nj >> nk;
    in >> min_box[0] >> min_box[1] >> min_box[2];
    in >> max_box[0] >> max_box[1] >> max_box[2];
    in >> spacing[0] >> spacing[1] >> spacing[2];
    const int64_t n = (int64_t)ni * nj * nk;
    SDFC_REQUIRE(!in.fail() && n > 0 && n <= capacity, "sdfc_sdf_load_text: malformed header or buffer too small");
    for (int64_t i = 0; i < n; ++i) in >> grid[i];
    return in.fail() ? fail(SDFC_ERR_INVALID, "sdfc_sdf_load_text: truncated file") : SDFC_OK;
}

}  // extern "C"
