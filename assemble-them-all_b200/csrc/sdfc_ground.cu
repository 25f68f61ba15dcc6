// sdfc_ground.cu — ForceGroundContact on sm_100a: the second Force
// plug-in of the free-fall configuration (examples/test_joint_sim.py:36,54), on the same
// "scan every contact point, evaluate the active ones, fixed-order reduction" skeleton as
// the SDF contact path.
//
//   k_ground<DERIV>  collision_detection_ground_object (CollisionDetection.cpp:37-68, mesh
//                    branch: d = ng . (E_0i xi - xg), active iff d <= 0) fused with
//                    ForceGroundContact::computeForce[WithDerivative]
//                    (Force/ForceGroundContact.cpp:106-242): normal + damping force, static /
//                    dynamic Coulomb branch, and the 6x6 body-diagonal blocks of Km and Dm.
//
// The reference forms everything with world-frame 3x3 products (R^T N R, R^T T R, ...).
// Here the ground normal is rotated into the body frame once per (env, ground),
// nb = R^T ng, which collapses those products to rank-one terms:
//   RNR = nb nb^T,  B = R^T T R = I - nb nb^T,  g = G phi = w x xi + v,  u = G^T nb
//   fm  += s u,  s = -kn d - c (nb.g)                                   (.cpp:127-129)
//   Km  -= G^T ( kn [ d [nb]x - nb (nb x xi)^T , nb nb^T ] + c [ (nb.g)[nb]x - nb (nb x g)^T , 0 ] )
//   Dm  -= c u u^T                                                      (.cpp:176-189)
//   static:  fm += G^T(-kt a), Dm += -kt G^T B G, Km += -kt G^T [ [a]x - B [g]x , 0 ],  a = B g
//   dynamic: t = a/|a|, A = (I - t t^T) B / |a|
//            fm += mukn d G^T t, Dm += mukn d G^T A G,
//            Km += mukn G^T ( [ d [t]x - d A [g]x , 0 ] + t u^T )       (.cpp:191-239)
// with G^T P = [ [xi]x P ; P ] for any 3 x k block P. The distance d that decides the active
// mask is computed in the reference's evaluation order with non-contractable intrinsics.
// Reductions are fixed-tree: thread -> warp shuffle -> warp order -> split order (last CTA).
#include "sdfc_host.h"

#include <algorithm>
#include <cstring>

namespace sdfc {

constexpr int kGroundThreads = 128;
constexpr int kGroundPointsPerCta = kGroundThreads * 32;
constexpr int kGroundOut = 78;  // fm[6], Km[36] column-major, Dm[36] column-major
constexpr double kConstEps = 1e-8;  // constants::eps (Utils.h:9)

struct GroundDev {
    const double *x, *y, *z;  // contact points of the body, SoA
    int n, body;
    double xg[3], ng[3];
    double kn, kt, mu, damping;
};

struct GroundBatchDev {
    const GroundDev* g;
    int n_grounds, n_bodies, n_envs, splits;
    double* partial;      // [n_envs][n_grounds][splits][78]
    double* out;          // [n_envs][n_grounds][78]
    int32_t* n_active;    // [n_envs][n_grounds]
    unsigned* ticket;     // [n_envs][n_grounds]
    const int* env_mode;  // see BatchDev::env_mode
    int env0, env_count;  // see BatchDev::env0
};

__device__ __forceinline__ void cross(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
// S = [a]x, row-major
__device__ __forceinline__ void skewm(const double* a, double* S) {
    S[0] = 0; S[1] = -a[2]; S[2] = a[1];
    S[3] = a[2]; S[4] = 0; S[5] = -a[0];
    S[6] = -a[1]; S[7] = a[0]; S[8] = 0;
}

// element t in [0, 12) of the top 3x4 of the column-major pose -> row-major R, p
__device__ __forceinline__ void stage_pose_g(const double* E, int n_bodies, int env, int body, double* R, double* p, int t) {
    const int r = t % 3, c = t / 3;
    const double v = __ldg(E + ((size_t)env * n_bodies + body) * 16 + r + 4 * c);
    if (c < 3) R[r * 3 + c] = v; else p[r] = v;
}

// acc += G^T P for a 3x6 row-major P, into a column-major 6x6 accumulator:
// rows 0..2 = xi x (columns of P), rows 3..5 = P
__device__ __forceinline__ void add_GtP(const double* xi, const double* P, double* acc) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const double p0 = P[c], p1 = P[6 + c], p2 = P[12 + c];
        acc[6 * c + 0] += xi[1] * p2 - xi[2] * p1;
        acc[6 * c + 1] += xi[2] * p0 - xi[0] * p2;
        acc[6 * c + 2] += xi[0] * p1 - xi[1] * p0;
        acc[6 * c + 3] += p0;
        acc[6 * c + 4] += p1;
        acc[6 * c + 5] += p2;
    }
}
// P(3x6) = M(3x3) G,  G = [ -[xi]x , I ]
__device__ __forceinline__ void MG(const double* M, const double* xi, double* P) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double m0 = M[3 * r], m1 = M[3 * r + 1], m2 = M[3 * r + 2];
        // m^T [xi]x = (m x xi)^T, so row r of -M [xi]x is (xi x m)^T
        P[6 * r + 0] = m2 * xi[1] - m1 * xi[2];
        P[6 * r + 1] = m0 * xi[2] - m2 * xi[0];
        P[6 * r + 2] = m1 * xi[0] - m0 * xi[1];
        P[6 * r + 3] = m0; P[6 * r + 4] = m1; P[6 * r + 5] = m2;
    }
}

template <bool DERIV>
__global__ void __launch_bounds__(kGroundThreads) k_ground(GroundBatchDev B, const double* __restrict__ E,
                                                           const double* __restrict__ phi) {
    const int split = blockIdx.x, gi = blockIdx.y, env = blockIdx.z + B.env0;
    if (B.env_mode != nullptr) {
        const int m = __ldg(B.env_mode + env);
        if (m == 0 || (m == 2) != DERIV) return;  // CTA-uniform
    }
    __shared__ GroundDev G;
    __shared__ double R[9], p[3], ph[6], nb[3];
    __shared__ double s_red[kGroundThreads / 32][kGroundOut];
    __shared__ int s_cnt[kGroundThreads / 32];
    __shared__ bool s_last;
    {
        const int* src = reinterpret_cast<const int*>(B.g + gi);
        int* dst = reinterpret_cast<int*>(&G);
        for (int i = threadIdx.x; i < (int)(sizeof(GroundDev) / sizeof(int)); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (threadIdx.x < 12) stage_pose_g(E, B.n_bodies, env, G.body, R, p, threadIdx.x);
    else if (threadIdx.x >= 32 && threadIdx.x < 38)
        ph[threadIdx.x - 32] = __ldg(phi + ((size_t)env * B.n_bodies + G.body) * 6 + (threadIdx.x - 32));
    __syncthreads();
    if (threadIdx.x < 3) nb[threadIdx.x] = R[threadIdx.x] * G.ng[0] + R[3 + threadIdx.x] * G.ng[1] + R[6 + threadIdx.x] * G.ng[2];
    __syncthreads();

    double acc[kGroundOut];  // the value-only instantiation touches acc[0..5] only; the rest is dead code
#pragma unroll
    for (int i = 0; i < (DERIV ? kGroundOut : 6); ++i) acc[i] = 0.0;
    int cnt = 0;
    const double kn = G.kn, kt = G.kt, mu = G.mu, cd = G.damping;
    const double* w = ph;
    const double* v = ph + 3;
    const int i0 = split * kGroundPointsPerCta;
    const int i1 = min(G.n, i0 + kGroundPointsPerCta);
    for (int i = i0 + threadIdx.x; i < i1; i += kGroundThreads) {
        double xi[3] = {__ldg(G.x + i), __ldg(G.y + i), __ldg(G.z + i)};
        double xw[3];
        to_world(R, p, xi[0], xi[1], xi[2], xw);
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(G.ng[0], __dsub_rn(xw[0], G.xg[0])),
                                             __dmul_rn(G.ng[1], __dsub_rn(xw[1], G.xg[1]))),
                                   __dmul_rn(G.ng[2], __dsub_rn(xw[2], G.xg[2])));
        if (!(d <= 0.0)) continue;
        ++cnt;
        double g[3], wx[3], u_top[3];
        cross(w, xi, wx);
#pragma unroll
        for (int a = 0; a < 3; ++a) g[a] = wx[a] + v[a];
        cross(xi, nb, u_top);  // u = G^T nb = [xi x nb ; nb]
        const double nbg = nb[0] * g[0] + nb[1] * g[1] + nb[2] * g[2];
        // force in the body frame: normal + damping, then friction
        double fb[3];
        const double s = -kn * d - cd * nbg;
#pragma unroll
        for (int a = 0; a < 3; ++a) fb[a] = s * nb[a];
        double P[18], MD[9];
        if (DERIV) {
            double nxi[3], nxg[3], Sn[9];
            cross(nb, xi, nxi);
            cross(nb, g, nxg);
            skewm(nb, Sn);
            const double c1 = kn * d + cd * nbg;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    P[6 * r + c] = -(c1 * Sn[3 * r + c] - nb[r] * (kn * nxi[c] + cd * nxg[c]));
                    P[6 * r + 3 + c] = -kn * nb[r] * nb[c];
                    MD[3 * r + c] = -cd * nb[r] * nb[c];
                }
        }
        if (!(mu < kConstEps)) {
            double a[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) a[k] = g[k] - nb[k] * nbg;
            const double an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            if (mu * fabs(kn * d) >= kt * an - kConstEps) {  // static (.cpp:137,197)
#pragma unroll
                for (int k = 0; k < 3; ++k) fb[k] += -kt * a[k];
                if (DERIV) {
                    double Sa[9], Sg[9];
                    skewm(a, Sa);
                    skewm(g, Sg);
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            // B [g]x = [g]x - nb (nb^T [g]x) = [g]x - nb (g x nb)^T ... nb^T [g]x = -(g x nb)^T = (nb x g)^T
                            const double nbSg = nb[0] * Sg[c] + nb[1] * Sg[3 + c] + nb[2] * Sg[6 + c];
                            const double BSg = Sg[3 * r + c] - nb[r] * nbSg;
                            P[6 * r + c] += -kt * (Sa[3 * r + c] - BSg);
                            MD[3 * r + c] += -kt * ((r == c ? 1.0 : 0.0) - nb[r] * nb[c]);
                        }
                }
            } else {  // dynamic (.cpp:144-149, 213-239)
                const double mukn = mu * kn;
                double t[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) t[k] = a[k] / an;
#pragma unroll
                for (int k = 0; k < 3; ++k) fb[k] += mukn * d * t[k];
                if (DERIV) {
                    // A = (I - t t^T)(I - nb nb^T) / |a|
                    double A[9], St[9], Sg[9];
                    const double tnb = t[0] * nb[0] + t[1] * nb[1] + t[2] * nb[2];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            A[3 * r + c] = ((r == c ? 1.0 : 0.0) - nb[r] * nb[c] - t[r] * t[c] + t[r] * tnb * nb[c]) / an;
                    skewm(t, St);
                    skewm(g, Sg);
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const double ASg = A[3 * r] * Sg[c] + A[3 * r + 1] * Sg[3 + c] + A[3 * r + 2] * Sg[6 + c];
                            P[6 * r + c] += mukn * (d * St[3 * r + c] - d * ASg + t[r] * u_top[c]);
                            P[6 * r + 3 + c] += mukn * t[r] * nb[c];
                            MD[3 * r + c] += mukn * d * A[3 * r + c];
                        }
                    }
                }
            }
        }
        // fm += G^T fb
        acc[0] += xi[1] * fb[2] - xi[2] * fb[1];
        acc[1] += xi[2] * fb[0] - xi[0] * fb[2];
        acc[2] += xi[0] * fb[1] - xi[1] * fb[0];
        acc[3] += fb[0]; acc[4] += fb[1]; acc[5] += fb[2];
        if constexpr (DERIV) {
            add_GtP(xi, P, acc + 6);
            double PD[18];
            MG(MD, xi, PD);
            add_GtP(xi, PD, acc + 42);
        }
    }
    // ---- fixed-order reduction: lanes -> warps -> (last CTA) splits
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NV = DERIV ? kGroundOut : 6;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[warp][i] = x;
    }
    {
        int c = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
        if (lane == 0) s_cnt[warp] = c;
    }
    __syncthreads();
    const size_t eg = (size_t)env * B.n_grounds + gi;
    double* part = B.partial + (eg * B.splits + split) * kGroundOut;
    if (threadIdx.x < NV) {
        double x = 0.0;
        for (int wv = 0; wv < kGroundThreads / 32; ++wv) x += s_red[wv][threadIdx.x];
        part[threadIdx.x] = x;
    }
    if (threadIdx.x == 0) {
        int c = 0;
        for (int wv = 0; wv < kGroundThreads / 32; ++wv) c += s_cnt[wv];
        if (c) atomicAdd(B.n_active + eg, c);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned tk = atomicAdd(B.ticket + eg, 1u);
        s_last = (tk == (unsigned)B.splits - 1u);
        if (s_last) B.ticket[eg] = 0u;  // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < kGroundOut) {
        double x = 0.0;
        if (threadIdx.x < NV)
            for (int sp = 0; sp < B.splits; ++sp) x += __ldcg(B.partial + (eg * B.splits + sp) * kGroundOut + threadIdx.x);
        B.out[eg * kGroundOut + threadIdx.x] = x;  // value-only launches zero the Km / Dm slots
    }
}

// out [env][ground][78] -> += into the per-environment system arrays (column-major)
__global__ void k_ground_scatter(GroundBatchDev B, int deriv, double* __restrict__ fm, double* __restrict__ Km,
                                 double* __restrict__ Dm) {
    const int env = blockIdx.x + B.env0;
    const int nm = 6 * B.n_bodies;
    for (int gi = 0; gi < B.n_grounds; ++gi) {  // serial over grounds: two grounds may share a body
        const double* o = B.out + ((size_t)env * B.n_grounds + gi) * kGroundOut;
        const int b0 = 6 * B.g[gi].body;
        for (int i = threadIdx.x; i < kGroundOut; i += blockDim.x) {
            if (i < 6) fm[(size_t)env * nm + b0 + i] += o[i];
            else if (deriv) {
                const int j = (i - 6) % 36, r = j % 6, c = j / 6;
                double* dst = (i < 42) ? Km : Dm;
                if (dst) dst[(size_t)env * nm * nm + (b0 + r) + (size_t)nm * (b0 + c)] += o[i];
            }
        }
        __syncthreads();
    }
}

}  // namespace sdfc

using namespace sdfc;
using sdfc_host::dev_alloc;
using sdfc_host::fail;

struct sdfc_ground {
    sdfc_context* ctx = nullptr;
    sdfc_points* pts = nullptr;
    double E_g[16];
    double kn = 1, kt = 0, mu = 0, damping = 0;
    uint64_t version = 1;
    sdfc_ground_set* single = nullptr;
    double* d_in = nullptr;  // E[16] + phi[6]
};

struct sdfc_ground_set {  // declared in sdfc_host.h
    sdfc_context* ctx = nullptr;
    int n_envs = 0, n_bodies = 0;
    std::vector<sdfc_ground*> grounds;
    std::vector<int32_t> body;
    std::vector<uint64_t> seen;
    GroundBatchDev dev{};
    GroundDev* d_g = nullptr;
    double *d_partial = nullptr, *d_out = nullptr;
    int32_t* d_n_active = nullptr;
    unsigned* d_ticket = nullptr;
};

namespace sdfc_host {

void ground_set_destroy(sdfc_ground_set* s) {
    if (!s) return;
    cudaFree(s->d_g); cudaFree(s->d_partial); cudaFree(s->d_out); cudaFree(s->d_n_active); cudaFree(s->d_ticket);
    delete s;
}

sdfc_status ground_set_create(sdfc_context* ctx, int n_envs, int n_bodies, int n, sdfc_ground* const* grounds,
                              const int32_t* body, sdfc_ground_set** out) {
    *out = nullptr;
    sdfc_ground_set* s = new sdfc_ground_set();
    s->ctx = ctx; s->n_envs = n_envs; s->n_bodies = n_bodies;
    s->grounds.assign(grounds, grounds + n);
    s->body.assign(body, body + n);
    s->seen.assign(n, 0);
    int max_n = 1;
    for (int i = 0; i < n; ++i) max_n = std::max(max_n, grounds[i]->pts->n);
    const int splits = (max_n + kGroundPointsPerCta - 1) / kGroundPointsPerCta;
    sdfc_status st = SDFC_OK;
    auto A = [&](sdfc_status x) { if (st == SDFC_OK) st = x; };
    const size_t eg = (size_t)n_envs * n;
    A(dev_alloc(&s->d_g, (size_t)n));
    A(dev_alloc(&s->d_partial, eg * splits * kGroundOut));
    A(dev_alloc(&s->d_out, eg * kGroundOut));
    A(dev_alloc(&s->d_n_active, eg));
    A(dev_alloc(&s->d_ticket, eg));
    // null-stream memset: synchronise before the (non-blocking) context stream launches anything that reads it
    if (st == SDFC_OK && (cudaMemset(s->d_ticket, 0, sizeof(unsigned) * eg) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess))
        st = fail(SDFC_ERR_CUDA, "ground_set_create: cudaMemset failed");
    if (st != SDFC_OK) { ground_set_destroy(s); return st; }
    GroundBatchDev& D = s->dev;
    D.g = s->d_g; D.n_grounds = n; D.n_bodies = n_bodies; D.n_envs = n_envs; D.splits = splits;
    D.partial = s->d_partial; D.out = s->d_out; D.n_active = s->d_n_active; D.ticket = s->d_ticket; D.env_mode = nullptr;
    *out = s;
    return SDFC_OK;
}

static sdfc_status ground_set_refresh(sdfc_ground_set* s) {
    bool stale = false;
    for (size_t i = 0; i < s->grounds.size(); ++i) stale |= s->seen[i] != s->grounds[i]->version;
    if (!stale) return SDFC_OK;
    std::vector<GroundDev> h(s->grounds.size());
    for (size_t i = 0; i < h.size(); ++i) {
        const sdfc_ground* g = s->grounds[i];
        GroundDev& d = h[i];
        const int n = g->pts->n;
        d.x = g->pts->d_s; d.y = g->pts->d_s + n; d.z = g->pts->d_s + 2 * (size_t)n;
        d.n = n; d.body = s->body[i];
        for (int a = 0; a < 3; ++a) { d.xg[a] = g->E_g[12 + a]; d.ng[a] = g->E_g[8 + a]; }  // column-major: col 3, col 2
        d.kn = g->kn; d.kt = g->kt; d.mu = g->mu; d.damping = g->damping;
        s->seen[i] = g->version;
    }
    // ordered after earlier launches on the stream that may still read the old descriptors
    SDFC_CUDA(cudaMemcpyAsync(s->d_g, h.data(), sizeof(GroundDev) * h.size(), cudaMemcpyHostToDevice, s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status ground_set_prepare(sdfc_ground_set* s) { return ground_set_refresh(s); }

// deriv / env_mode_dev as in batch_run. Results stay in the set's buffers.
sdfc_status ground_set_run(sdfc_ground_set* s, const double* E_dev, const double* phi_dev, int deriv,
                           const int* env_mode_dev, int env0, int env_count, cudaStream_t stream_override,
                           bool counts_zeroed) {
    sdfc_status st = ground_set_refresh(s);
    if (st != SDFC_OK) return st;
    cudaStream_t stream = stream_override ? stream_override : s->ctx->stream;
    GroundBatchDev D = s->dev;
    D.env_mode = env_mode_dev;
    if (env_count < 0) { env0 = 0; env_count = s->n_envs; }
    D.env0 = env0; D.env_count = env_count;
    if (!counts_zeroed)
        SDFC_CUDA(cudaMemsetAsync(s->d_n_active + (size_t)env0 * D.n_grounds, 0, sizeof(int32_t) * (size_t)env_count * D.n_grounds, stream));
    dim3 grid(D.splits, D.n_grounds, env_count);
    if (env_mode_dev != nullptr) {
        k_ground<false><<<grid, kGroundThreads, 0, stream>>>(D, E_dev, phi_dev);
        k_ground<true><<<grid, kGroundThreads, 0, stream>>>(D, E_dev, phi_dev);
        s->ctx->launches += 2;
    } else {
        if (deriv) k_ground<true><<<grid, kGroundThreads, 0, stream>>>(D, E_dev, phi_dev);
        else k_ground<false><<<grid, kGroundThreads, 0, stream>>>(D, E_dev, phi_dev);
        s->ctx->launches += 1;
    }
    SDFC_CUDA(cudaGetLastError());
    return SDFC_OK;
}

sdfc_status ground_set_scatter(sdfc_ground_set* s, int deriv, double* fm_dev, double* Km_dev, double* Dm_dev) {
    GroundBatchDev D = s->dev;
    D.env0 = 0; D.env_count = s->n_envs;
    k_ground_scatter<<<s->n_envs, 96, 0, s->ctx->stream>>>(D, deriv, fm_dev, Km_dev, Dm_dev);
    s->ctx->launches += 1;
    SDFC_CUDA(cudaGetLastError());
    return SDFC_OK;
}

const double* ground_set_out(const sdfc_ground_set* s) { return s->d_out; }
int32_t* ground_set_n_active(const sdfc_ground_set* s) { return s->d_n_active; }
int ground_set_count(const sdfc_ground_set* s) { return (int)s->grounds.size(); }
const int32_t* ground_set_bodies(const sdfc_ground_set* s) { return s->body.data(); }

}  // namespace sdfc_host

extern "C" {

sdfc_status sdfc_ground_create(sdfc_context* ctx, sdfc_points* contact_points, const double E_g[16], double kn, double kt,
                               double mu, double damping, sdfc_ground** out) {
    SDFC_REQUIRE(ctx && contact_points && E_g && out, "sdfc_ground_create: NULL argument");
    SDFC_REQUIRE(contact_points->ctx == ctx, "sdfc_ground_create: contact points belong to another context");
    *out = nullptr;
    SDFC_CUDA(cudaSetDevice(ctx->device));
    sdfc_ground* g = new sdfc_ground();
    g->ctx = ctx; g->pts = contact_points;
    std::memcpy(g->E_g, E_g, sizeof(g->E_g));
    g->kn = kn; g->kt = kt; g->mu = mu; g->damping = damping;
    sdfc_status st = dev_alloc(&g->d_in, (size_t)22);
    if (st != SDFC_OK) { delete g; return st; }
    *out = g;
    return SDFC_OK;
}
sdfc_status sdfc_ground_set_transform(sdfc_ground* g, const double E_g[16]) {
    SDFC_REQUIRE(g && E_g, "sdfc_ground_set_transform: NULL argument");
    if (std::memcmp(g->E_g, E_g, sizeof(g->E_g)) == 0) return SDFC_OK;   // unchanged: no re-upload
    std::memcpy(g->E_g, E_g, sizeof(g->E_g)); g->version++;
    return SDFC_OK;
}
sdfc_status sdfc_ground_set_stiffness(sdfc_ground* g, double kn, double kt) {
    SDFC_REQUIRE(g != nullptr, "sdfc_ground_set_stiffness: ground is NULL");
    if (g->kn == kn && g->kt == kt) return SDFC_OK;
    g->kn = kn; g->kt = kt; g->version++;
    return SDFC_OK;
}
sdfc_status sdfc_ground_set_friction(sdfc_ground* g, double mu) {
    SDFC_REQUIRE(g != nullptr, "sdfc_ground_set_friction: ground is NULL");
    if (g->mu == mu) return SDFC_OK;
    g->mu = mu; g->version++;
    return SDFC_OK;
}
sdfc_status sdfc_ground_set_damping(sdfc_ground* g, double damping) {
    SDFC_REQUIRE(g != nullptr, "sdfc_ground_set_damping: ground is NULL");
    if (g->damping == damping) return SDFC_OK;
    g->damping = damping; g->version++;
    return SDFC_OK;
}
sdfc_status sdfc_ground_destroy(sdfc_ground* g) {
    if (!g) return SDFC_OK;
    cudaSetDevice(g->ctx->device);
    sdfc_host::ground_set_destroy(g->single);
    cudaFree(g->d_in);
    delete g;
    return SDFC_OK;
}

static sdfc_status ground_eval(sdfc_ground* g, const double E[16], const double phi[6], int deriv, double fm_sub[6],
                               double Km_sub[36], double Dm_sub[36], int32_t* n_active) {
    SDFC_CUDA(cudaSetDevice(g->ctx->device));
    if (!g->single) {
        const int32_t body0 = 0;
        sdfc_ground* self = g;
        sdfc_status st = sdfc_host::ground_set_create(g->ctx, 1, 1, 1, &self, &body0, &g->single);
        if (st != SDFC_OK) return st;
    }
    cudaStream_t s = g->ctx->stream;
    double in[22];
    std::memcpy(in, E, 16 * sizeof(double));
    std::memcpy(in + 16, phi, 6 * sizeof(double));
    SDFC_CUDA(cudaMemcpyAsync(g->d_in, in, sizeof(in), cudaMemcpyHostToDevice, s));
    SDFC_CUDA(cudaStreamSynchronize(s));  // `in` is pageable stack memory
    sdfc_status st = sdfc_host::ground_set_run(g->single, g->d_in, g->d_in + 16, deriv, nullptr, 0, -1);
    if (st != SDFC_OK) return st;
    double out[kGroundOut];
    int32_t na = 0;
    SDFC_CUDA(cudaMemcpyAsync(out, g->single->d_out, sizeof(out), cudaMemcpyDeviceToHost, s));
    SDFC_CUDA(cudaMemcpyAsync(&na, g->single->d_n_active, sizeof(na), cudaMemcpyDeviceToHost, s));
    SDFC_CUDA(cudaStreamSynchronize(s));
    std::memcpy(fm_sub, out, 6 * sizeof(double));
    if (deriv) {
        std::memcpy(Km_sub, out + 6, 36 * sizeof(double));
        std::memcpy(Dm_sub, out + 42, 36 * sizeof(double));
    }
    if (n_active) *n_active = na;
    return SDFC_OK;
}

sdfc_status sdfc_ground_force(sdfc_ground* g, const double E[16], const double phi[6], double fm_sub[6], int32_t* n_active) {
    SDFC_REQUIRE(g && E && phi && fm_sub, "sdfc_ground_force: NULL argument");
    return ground_eval(g, E, phi, 0, fm_sub, nullptr, nullptr, n_active);
}
sdfc_status sdfc_ground_force_deriv(sdfc_ground* g, const double E[16], const double phi[6], double fm_sub[6],
                                    double Km_sub[36], double Dm_sub[36], int32_t* n_active) {
    SDFC_REQUIRE(g && E && phi && fm_sub && Km_sub && Dm_sub, "sdfc_ground_force_deriv: NULL argument");
    return ground_eval(g, E, phi, 1, fm_sub, Km_sub, Dm_sub, n_active);
}

}  // extern "C"
