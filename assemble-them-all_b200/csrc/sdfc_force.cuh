// sdfc_force.cuh — the force / Jacobian kernel (LOOP D) as templates, so that each
// (DERIV, ANYFRIC, ANYDAMP) variant compiles in its own translation unit
// (sdfc_force_inst.cu is compiled 8 times with -DSDFC_FORCE_VARIANT=0..7).
#pragma once
#include "sdfc_kernels.cuh"

namespace sdfc {

// --------------------------------------------------------------------------
// register transpose-reduce across the 32 lanes of a warp
// --------------------------------------------------------------------------
// Every lane holds N values; afterwards lane L holds tr_out(N) values, each the
// sum over all 32 lanes of one original entry (entry = tr_index<N>(L, j)). Each
// step exchanges half of the remaining values with the lane at distance MASK, so
// the cost is ~N shuffles instead of 5N for N independent butterfly sums.
template <int N, int MASK>
struct TransposeReduce {
    static __device__ __forceinline__ void run(double* v, unsigned lane) {
        constexpr int H = (N + 1) / 2;
        const bool up = (lane & MASK) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double lo = v[i];
            const double hi = (H + i < N) ? v[H + i] : 0.0;
            const double send = up ? lo : hi;
            const double keep = up ? hi : lo;
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
        }
        TransposeReduce<H, MASK / 2>::run(v, lane);
    }
};
template <int N>
struct TransposeReduce<N, 0> {
    static __device__ __forceinline__ void run(double*, unsigned) {}
};

// Original entry held by (lane, j) after TransposeReduce<N,16>, or -1 for padding.
template <int N>
__host__ __device__ inline int tr_index(int lane, int j) {
    int sizes[6];
    sizes[0] = N;
    for (int s = 0; s < 5; ++s) sizes[s + 1] = (sizes[s] + 1) / 2;
    int rel = j;
    if (rel >= sizes[5]) return -1;
    for (int s = 4; s >= 0; --s) {
        if (lane & (16 >> s)) rel += sizes[s + 1];
        if (rel >= sizes[s]) return -1;
    }
    return rel;
}

// Fixed-order sum of the split partials of one (environment, pair) into its blocks fm_sub[12], Km_sub[144],
// Dm_sub[144] (column-major), undoing the slot layout of the warp transpose-reduce. `tid` of `nthreads`
// cooperating threads (k_reduce_splits: a CTA; the stepper's round kernel: the environment's warp).
__device__ __forceinline__ void reduce_splits_one(const BatchDev& B, int env, int pair, int deriv, int dm, int tid, int nthreads,
                                                  double* fm_sub, double* Km_sub, double* Dm_sub) {
    const size_t ep = (size_t)env * B.n_pairs + pair;
    const bool empty = B.n_active[ep] == 0;
    const bool nf = deriv && !dm;  // layout of contact_eval_nf: 12 Km columns + fm
    const int npart = nf ? kPartialNF : (deriv ? (dm ? kPartialDerivDM : kPartialDeriv) : kPartialValue);
    const double* part = B.partial + ep * B.splits * kPartialDerivDM;  // fixed slot stride, see k_force
    for (int slot = tid; slot < npart; slot += nthreads) {
        double s = 0.0;
        if (!empty)
            for (int sp = 0; sp < B.splits; ++sp) s += part[(size_t)sp * kPartialDerivDM + slot];
        const int lane = slot & 31;
        if (nf) {
            const int j = slot >> 5, e = tr_index<12>(lane, 0);
            if (e >= 0) {
                if (j < 12) Km_sub[ep * 144 + e + 12 * j] = s;
                else fm_sub[ep * 12 + e] = s;
            }
        } else if (!deriv) {
            const int e = tr_index<12>(lane, slot >> 5);
            if (e >= 0) fm_sub[ep * 12 + e] = s;
        } else if (slot < kPartialDeriv) {
            const int b = (slot >> 5) / kOutK, j = (slot >> 5) % kOutK;
            const int e = tr_index<39>(lane, j);
            if (e >= 36) fm_sub[ep * 12 + 3 * b + (e - 36)] = s;
            else if (e >= 0) {
                const int r = e / 6 + 6 * (b >> 1), c = e % 6 + 6 * (b & 1);
                Km_sub[ep * 144 + r + 12 * c] = s;
            }
        } else {
            const int q = (slot - kPartialDeriv) >> 5;
            const int b = q / kOutD, j = q % kOutD;
            const int e = tr_index<36>(lane, j);
            if (e >= 0) {
                const int r = e / 6 + 6 * (b >> 1), c = e % 6 + 6 * (b & 1);
                Dm_sub[ep * 144 + r + 12 * c] = s;
            }
        }
    }
    if (deriv && !dm && Dm_sub != nullptr)
        for (int i = tid; i < 144; i += nthreads) Dm_sub[ep * 144 + i] = 0.0;
}

// Sink handed to contact_eval: reduces each emitted block across the warp at once.
template <bool DERIV, bool DM>
struct WarpSink {
    double accK[DERIV ? 4 : 1][kOutK];
    double accD[(DERIV && DM) ? 4 : 1][kOutD];
    double accF[kOutF];
    unsigned lane;
    bool valid;  // false: this lane re-evaluates a valid contact but contributes zeros

    __device__ __forceinline__ void init(unsigned l) {
        lane = l;
        valid = false;
#pragma unroll
        for (int b = 0; b < (DERIV ? 4 : 1); ++b)
#pragma unroll
            for (int j = 0; j < kOutK; ++j) accK[b][j] = 0.0;
#pragma unroll
        for (int b = 0; b < ((DERIV && DM) ? 4 : 1); ++b)
#pragma unroll
            for (int j = 0; j < kOutD; ++j) accD[b][j] = 0.0;
#pragma unroll
        for (int j = 0; j < kOutF; ++j) accF[j] = 0.0;
    }
    __device__ __forceinline__ void emitF(const double* fm12) {
        double v[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = valid ? fm12[i] : 0.0;
        TransposeReduce<12, 16>::run(v, lane);
#pragma unroll
        for (int j = 0; j < kOutF; ++j) accF[j] += v[j];
    }
    template <int B>
    __device__ __forceinline__ void emitK(const double* k36, const double* fm3) {
        double v[39];
#pragma unroll
        for (int i = 0; i < 36; ++i) v[i] = valid ? k36[i] : 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) v[36 + i] = valid ? fm3[i] : 0.0;
        TransposeReduce<39, 16>::run(v, lane);
#pragma unroll
        for (int j = 0; j < kOutK; ++j) accK[DERIV ? B : 0][j] += v[j];
    }
    template <int B>
    __device__ __forceinline__ void emitD(const double* d36) {
        double v[36];
#pragma unroll
        for (int i = 0; i < 36; ++i) v[i] = valid ? d36[i] : 0.0;
        TransposeReduce<36, 16>::run(v, lane);
#pragma unroll
        for (int j = 0; j < kOutD; ++j) accD[(DERIV && DM) ? B : 0][j] += v[j];
    }
};

// Sink of contact_eval_nf: one 12-value column at a time.
struct ColSink {
    double acc[kNfCols];
    unsigned lane;
    bool valid;
    __device__ __forceinline__ void init(unsigned l) {
        lane = l;
        valid = false;
#pragma unroll
        for (int j = 0; j < kNfCols; ++j) acc[j] = 0.0;
    }
    __device__ __forceinline__ void emit(int j, const double* c12) {
        // lanes beyond the last contact evaluate a duplicate with kn = 0 (force_rounds_nf): every
        // output of contact_eval_nf is proportional to kn, so they contribute exact zeros without
        // 24 selects per column here
        double v[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = c12[i];
        TransposeReduce<12, 16>::run(v, lane);
        acc[j] += v[0];
    }
};

// Pose of (env, body) from column-major E[env][body][16] into row-major R, p.
__device__ __forceinline__ void stage_pose(const double* E, int n_bodies, int env, int body, double* R, double* p,
                                           int t) {
    // t in [0, 12): element (r, c) of the top 3x4
    const int r = t % 3, c = t / 3;
    const double v = __ldg(E + ((size_t)env * n_bodies + body) * 16 + r + 4 * c);
    if (c < 3) R[r * 3 + c] = v; else p[r] = v;
}

// --------------------------------------------------------------------------
// force
// --------------------------------------------------------------------------
struct ForceShared {
    PairDev pd;
    PairPose pose;
    PairPoseNF nf;
    int warp_tot[kNumWarps];
    int ids[kForceThreads * 32];
};

template <bool F32, bool DERIV, bool FRIC, bool DAMP, class Sink>
__device__ __forceinline__ void force_rounds(const ForceShared& S, int total, Sink& sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = warp * 32; base < total; base += kForceThreads) {
        const int c = base + lane;
        sink.valid = c < total;
        const int id = S.ids[sink.valid ? c : base];
        const double xi1[3] = {__ldg(S.pd.sx + id), __ldg(S.pd.sy + id), __ldg(S.pd.sz + id)};
        contact_eval<F32, DERIV, FRIC, DAMP>(S.pd.grid, S.pose, S.pd.prm, xi1, sink);
    }
}

// Sink of contact_eval_nf2: 10 reductions of 12 values per round (Km columns 0..2, 6..8, the three columns of Y, fm).
struct Nf2Sink {
    double acc[kNf2Slots];
    unsigned lane;
    __device__ __forceinline__ void init(unsigned l) {
        lane = l;
#pragma unroll
        for (int j = 0; j < kNf2Slots; ++j) acc[j] = 0.0;
    }
    __device__ __forceinline__ void emit(int j, const double* c12) {
        // lanes beyond the last contact evaluate a duplicate with kn = 0: every output is proportional to kn
        double v[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = c12[i];
        TransposeReduce<12, 16>::run(v, lane);
        acc[j] += v[0];
    }
};

template <bool F32>
__device__ __forceinline__ void force_rounds_nf(const ForceShared& S, int total, Nf2Sink& sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = warp * 32; base < total; base += kForceThreads) {
        const int c = base + lane;
        const bool valid = c < total;
        const int id = S.ids[valid ? c : base];
        const double xi1[3] = {__ldg(S.pd.sx + id), __ldg(S.pd.sy + id), __ldg(S.pd.sz + id)};
        contact_eval_nf2<F32>(S.pd.grid, S.nf, valid ? S.pd.prm.kn : 0.0, xi1, sink);
    }
}

#ifndef SDFC_FORCE_MIN_BLOCKS
#define SDFC_FORCE_MIN_BLOCKS 2
#endif
#ifndef SDFC_FORCE_NF_MIN_BLOCKS
#define SDFC_FORCE_NF_MIN_BLOCKS 3  // planner fast path: 168 registers, 12 warps / SM
#endif
template <bool DERIV, bool ANYFRIC, bool ANYDAMP>
__global__ void __launch_bounds__(kForceThreads, (DERIV && !ANYFRIC && !ANYDAMP) ? SDFC_FORCE_NF_MIN_BLOCKS
                                                                                : SDFC_FORCE_MIN_BLOCKS) k_force(BatchDev B, const double* __restrict__ E,
                                                         const double* __restrict__ phi) {
    constexpr bool DM = ANYFRIC || ANYDAMP;
    constexpr bool NF = DERIV && !DM;  // planner fast path (contact_eval_nf)
    constexpr int NPART = NF ? kPartialNF : (DERIV ? (DM ? kPartialDerivDM : kPartialDeriv) : kPartialValue);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ForceShared& S = *reinterpret_cast<ForceShared*>(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw + sizeof(ForceShared));  // [kNumWarps][NPART]

    const int env = blockIdx.y + B.env0;
    if (B.env_mode != nullptr && __ldg(B.env_mode + env) != (DERIV ? 2 : 1)) return;  // CTA-uniform
    const int pair = blockIdx.x / B.splits, split = blockIdx.x % B.splits;
    if (__ldg(B.n_active + (size_t)env * B.n_pairs + pair) == 0) return;  // CTA-uniform
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const int* src = reinterpret_cast<const int*>(B.pairs + pair);
        int* dst = reinterpret_cast<int*>(&S.pd);
        for (int i = tid; i < (int)(sizeof(PairDev) / sizeof(int)); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    // 36 threads stage the two poses and twists (any CTA width from one warp pair up)
    if (tid < 12) stage_pose(E, B.n_bodies, env, S.pd.body_a, S.pose.R1, S.pose.p1, tid);
    else if (tid < 24) stage_pose(E, B.n_bodies, env, S.pd.body_b, S.pose.R2, S.pose.p2, tid - 12);
    else if (tid < 30) S.pose.phi1[tid - 24] = __ldg(phi + ((size_t)env * B.n_bodies + S.pd.body_a) * 6 + tid - 24);
    else if (tid < 36) S.pose.phi2[tid - 30] = __ldg(phi + ((size_t)env * B.n_bodies + S.pd.body_b) * 6 + tid - 30);

    if constexpr (NF) {
        __syncthreads();
        if (tid < 9) {
            const int r = tid / 3, c = tid % 3;  // R21 = R2^T R1
            S.nf.R21[tid] = S.pose.R2[r] * S.pose.R1[c] + S.pose.R2[3 + r] * S.pose.R1[3 + c] + S.pose.R2[6 + r] * S.pose.R1[6 + c];
        } else if (tid >= 9 && tid < 18) {
            S.nf.R1[tid - 9] = S.pose.R1[tid - 9]; S.nf.R2[tid - 9] = S.pose.R2[tid - 9];
        } else if (tid >= 18 && tid < 21) {
            S.nf.p1[tid - 18] = S.pose.p1[tid - 18]; S.nf.p2[tid - 18] = S.pose.p2[tid - 18];
        }
    }
    WarpSink<DERIV && !NF, DM> sink;
    Nf2Sink csink;
    if constexpr (NF) csink.init(lane); else sink.init(lane);

    const int wps = (S.pd.n_words + B.splits - 1) / B.splits;
    const int w0 = split * wps, w1 = min(S.pd.n_words, w0 + wps);
    const uint32_t* mask_row = B.mask + (size_t)env * B.words_per_env + S.pd.word_offset;
    const bool f32 = S.pd.grid.f32();
    const bool fric = ANYFRIC && (S.pd.prm.mu > kEps);        // .cpp:200,352
    const bool damp = ANYDAMP && (S.pd.prm.damping != 0.0);

    for (int tw = w0; tw < w1; tw += kForceThreads) {
        __syncthreads();  // poses staged / previous tile's ids consumed
        uint32_t word = (tw + tid < w1) ? __ldg(mask_row + tw + tid) : 0u;
        const int cnt = __popc(word);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) S.warp_tot[warp] = incl;
        __syncthreads();
        int off = incl - cnt, total = 0;
#pragma unroll
        for (int w = 0; w < kNumWarps; ++w) {
            const int t = S.warp_tot[w];
            if (w < warp) off += t;
            total += t;
        }
        while (word) {
            const int b = __ffs(word) - 1;
            S.ids[off++] = (tw + tid) * 32 + b;
            word &= word - 1;
        }
        __syncthreads();
        if (total == 0) continue;
        if constexpr (NF) {
            if (f32) force_rounds_nf<true>(S, total, csink); else force_rounds_nf<false>(S, total, csink);
            continue;
        }
        // CTA-uniform variant dispatch (template flags prune what cannot occur)
        if (f32) {
            if (fric) { if (damp) force_rounds<true, DERIV, ANYFRIC, ANYDAMP>(S, total, sink); else force_rounds<true, DERIV, ANYFRIC, false>(S, total, sink); }
            else      { if (damp) force_rounds<true, DERIV, false, ANYDAMP>(S, total, sink);  else force_rounds<true, DERIV, false, false>(S, total, sink); }
        } else {
            if (fric) { if (damp) force_rounds<false, DERIV, ANYFRIC, ANYDAMP>(S, total, sink); else force_rounds<false, DERIV, ANYFRIC, false>(S, total, sink); }
            else      { if (damp) force_rounds<false, DERIV, false, ANYDAMP>(S, total, sink);  else force_rounds<false, DERIV, false, false>(S, total, sink); }
        }
    }

    // cross-warp combine in fixed order
    double* mine = red + (size_t)warp * NPART;
    if constexpr (NF) {
#pragma unroll
        for (int j = 0; j < kNf2Slots; ++j) mine[j * 32 + lane] = csink.acc[j];
    } else if (DERIV) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int j = 0; j < kOutK; ++j) mine[(b * kOutK + j) * 32 + lane] = sink.accK[DERIV ? b : 0][j];
        if (DM) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int j = 0; j < kOutD; ++j)
                    mine[kPartialDeriv + (b * kOutD + j) * 32 + lane] = sink.accD[(DERIV && DM) ? b : 0][j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < kOutF; ++j) mine[j * 32 + lane] = sink.accF[j];
    }
    __syncthreads();
    // fixed slot stride (the largest layout): value-only and derivative environments share the
    // buffer within one evaluation round of the stepper, so the stride must not depend on the variant
    double* out = B.partial + (((size_t)env * B.n_pairs + pair) * B.splits + split) * kPartialDerivDM;
    if constexpr (NF) {
        // 10 reduced slots -> the 13-column layout of the partials: columns 3..5 = (sum Y) R21, 9..11 = -(sum Y)
        if (tid < 32) {
            double a[kNf2Slots];
#pragma unroll
            for (int j = 0; j < kNf2Slots; ++j) {
                double s = red[j * 32 + tid];
#pragma unroll
                for (int w = 1; w < kNumWarps; ++w) s += red[(size_t)w * NPART + j * 32 + tid];
                a[j] = s;
            }
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                out[m * 32 + tid] = a[m];
                out[(6 + m) * 32 + tid] = a[3 + m];
                out[(3 + m) * 32 + tid] = a[6] * S.nf.R21[m] + a[7] * S.nf.R21[3 + m] + a[8] * S.nf.R21[6 + m];
                out[(9 + m) * 32 + tid] = -a[6 + m];
            }
            out[12 * 32 + tid] = a[9];
        }
        return;
    }
    for (int i = tid; i < NPART; i += kForceThreads) {
        double s = red[i];
#pragma unroll
        for (int w = 1; w < kNumWarps; ++w) s += red[(size_t)w * NPART + i];
        out[i] = s;
    }
}

template <bool DERIV, bool ANYFRIC, bool ANYDAMP>
void launch_force_t(const BatchDev& B, const double* E, const double* phi, cudaStream_t stream) {
    constexpr bool DM = ANYFRIC || ANYDAMP;
    {
    constexpr int NPART = (DERIV && !DM) ? kPartialNF : (DERIV ? (DM ? kPartialDerivDM : kPartialDeriv) : kPartialValue);
    const size_t smem = sizeof(ForceShared) + sizeof(double) * kNumWarps * NPART;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_force<DERIV, ANYFRIC, ANYDAMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid(B.n_pairs * B.splits, B.env_count);
    k_force<DERIV, ANYFRIC, ANYDAMP><<<grid, kForceThreads, smem, stream>>>(B, E, phi);
    }
}


}  // namespace sdfc
