// sdfc_force.cuh — the force / Jacobian kernel (LOOP D) as templates, so that each
// (DERIV, ANYFRIC, ANYDAMP) variant compiles in its own translation unit
// (sdfc_force_inst.cu is compiled 8 times with -DSDFC_FORCE_VARIANT=0..7).
#pragma once
#include "sdfc_kernels.cuh"
#include <algorithm>
#include <type_traits>

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

// Fixed-order sum of the item partials of one (environment, pair) into its blocks fm_sub[12], Km_sub[144],
// Dm_sub[144] (column-major), undoing the slot layout of the warp transpose-reduce. The pair's items are
// consecutive in its plan (ForcePlanDev) and are summed in item order, so the result does not depend on which
// warp evaluated which item. `tid` of `nthreads` cooperating threads (k_reduce_items: a CTA; the stepper's
// round kernel: the environment's warp).
__device__ __forceinline__ void reduce_items_one(const BatchDev& B, int env, int pair, int deriv, int dm, int tid, int nthreads,
                                                 double* fm_sub, double* Km_sub, double* Dm_sub) {
    const size_t ep = (size_t)env * B.n_pairs + pair;
    const ForcePlanDev& P = B.plan[deriv ? 1 : 0];
    const int n_act = B.n_active[ep];
    const int per_item = 32 * P.hdr[3];
    const int n_items = (n_act + per_item - 1) / per_item;
    const bool nf = deriv && !dm;  // layout of the planner fast path: 13 columns x 2 halves x 12 entries
    const int npart = nf ? kNfCols * 12 : (deriv ? kPartialDerivDM : kPartialValue);
    const double* part = n_items ? P.partial + (size_t)P.pair_item0[ep] * P.stride : nullptr;
    for (int slot = tid; slot < npart; slot += nthreads) {
        if (nf) {
            // output (column j, entry e) = sum over the items, in item order, of its two half-sums
            const int j = slot / 12, e = slot - 12 * j;
            const double* src = part + j * kNfLanes + e;
            double s = 0.0;
            int it = 0;
            for (; it + 4 <= n_items; it += 4) {   // eight loads in flight, additions strictly in order
                double a[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a[2 * u] = __ldcg(src + (size_t)(it + u) * P.stride);
                    a[2 * u + 1] = __ldcg(src + (size_t)(it + u) * P.stride + 12);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) s += a[u];
            }
            for (; it < n_items; ++it) { s += __ldcg(src + (size_t)it * P.stride); s += __ldcg(src + (size_t)it * P.stride + 12); }
            if (j < 12) Km_sub[ep * 144 + e + 12 * j] = s;
            else fm_sub[ep * 12 + e] = s;
            continue;
        }
        // item order, eight loads in flight: the additions stay strictly sequential (same bits as a plain loop)
        double s = 0.0;
        int it = 0;
        for (; it + 8 <= n_items; it += 8) {
            double a[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = __ldcg(part + (size_t)(it + u) * P.stride + slot);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += a[u];
        }
        for (; it < n_items; ++it) s += __ldcg(part + (size_t)it * P.stride + slot);
        const int lane = slot & 31;
        if (!deriv) {
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

// Sink of contact_eval_nf: one 12-value column at a time, summed over the warp's 32 contacts through shared
// memory. Every lane stores its column into a 12 x 33 tile; lanes 0..23 then each add 16 contacts of ONE entry
// (entry r = lane % 12 over the contacts of half h = lane / 12, two chains of 8 in a fixed order) and keep the
// running sum of the item in the warp's acc block. 12 STS + 16 LDS + 16 DADD per column against ~120
// instructions (SHFL / FSEL / MOV / DADD) of the register transpose-reduce; the tiles are double-buffered, so
// one __syncwarp per column orders everything (a tile is rewritten two emits later, i.e. after the next sync).
constexpr int kNfTile = 12 * 33;
struct ColSink {
    double* tile;        // the warp's two transpose tiles
    double* acc;         // this lane's slot-0 accumulator: acc[slot * kNfLanes]
    const double* rd;    // first element this lane reads of tile 0
    int lane, rot, buf;
    __device__ __forceinline__ void init(double* warp_tiles, double* warp_acc, int l) {
        lane = l;
        buf = 0;
        tile = warp_tiles;
        const int r = l % 12, h = (l / 12) & 1;
        rd = warp_tiles + r * 33 + 16 * h;
        rot = 12 * h;    // the second half reads its 16 contacts rotated by 12: no bank shared with the first
        acc = warp_acc + l;
        if (l < kNfLanes)
#pragma unroll
            for (int j = 0; j < kNfSlots; ++j) acc[j * kNfLanes] = 0.0;
    }
    __device__ __forceinline__ void emit(int slot, const double* c12) {
        // lanes beyond the last contact evaluate a duplicate with kn = 0 (force_round): every output of
        // contact_eval_nf is proportional to kn, so they contribute exact zeros
        buf ^= kNfTile;   // the other tile
        double* t = tile + buf + lane;
#pragma unroll
        for (int i = 0; i < 12; ++i) t[i * 33] = c12[i];
        __syncwarp();
        if (lane < kNfLanes) {
            const double* b = rd + buf;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s0 += b[(i + rot) & 15];
                s1 += b[(i + 8 + rot) & 15];
            }
            acc[slot * kNfLanes] += s0 + s1;
        }
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
// force: a persistent kernel over the work list of the launch
// --------------------------------------------------------------------------
// The active contacts of an evaluation are spread very unevenly over the (environment, pair)s (none in most,
// hundreds in some). k_force_plan (sdfc_kernels.cu) cuts them into ITEMS of G x 32 consecutive contacts of one
// pair (mask order; G = 1 unless the list would overflow its capacity); the warps of this kernel take items
// from a device-side counter until none is left, so every warp evaluates full 32-contact rounds whatever the
// distribution is. An item's sums go to its own slot of the plan's partial buffer and the slots of a pair are
// added in item order afterwards (reduce_items_one): the result is independent of the schedule.
template <bool NF>
struct ForceWarpShared {
    PairDev pd;
    typename std::conditional<NF, PairPoseNF, PairPose>::type pose;
    double acc[NF ? kNfSlots * kNfLanes : 1];   // planner fast path: the running sums of the item (ColSink)
    double tile[NF ? 2 * kNfTile : 1];          // ... and its transpose tiles
};

// accumulators of one item -> its slot, then cleared
template <bool DERIV, bool DM>
__device__ __forceinline__ void flush_item(WarpSink<DERIV, DM>& sink, double* out, int lane) {
    if constexpr (DERIV) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int j = 0; j < kOutK; ++j) { out[(b * kOutK + j) * 32 + lane] = sink.accK[b][j]; sink.accK[b][j] = 0.0; }
        if constexpr (DM) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int j = 0; j < kOutD; ++j) { out[kPartialDeriv + (b * kOutD + j) * 32 + lane] = sink.accD[b][j]; sink.accD[b][j] = 0.0; }
        }
    } else {
#pragma unroll
        for (int j = 0; j < kOutF; ++j) { out[j * 32 + lane] = sink.accF[j]; sink.accF[j] = 0.0; }
    }
}
// the 10 summed slots of the item -> the 13 columns of its partial (kPartialNF layout), lane = (half, entry)
__device__ __forceinline__ void flush_item(ColSink& sink, const double* R21, double* out, int lane) {
    if (lane < kNfLanes) {
        double a[kNfSlots], col[kNfCols];
#pragma unroll
        for (int j = 0; j < kNfSlots; ++j) { a[j] = sink.acc[j * kNfLanes]; sink.acc[j * kNfLanes] = 0.0; }
        nf_slots_to_columns(a, R21, col);
#pragma unroll
        for (int j = 0; j < kNfCols; ++j) out[j * kNfLanes + lane] = col[j];
    }
}

// One round: contact `id` on this lane (`valid` = false: a duplicate that contributes zeros).
template <bool F32, bool DERIV, bool FRIC, bool DAMP, bool NF, class Sink>
__device__ __forceinline__ void force_round(const ForceWarpShared<NF>& S, int id, bool valid, Sink& sink) {
    const double xi1[3] = {__ldg(S.pd.sx + id), __ldg(S.pd.sy + id), __ldg(S.pd.sz + id)};
    if constexpr (NF) {
        // lanes past the last contact evaluate a duplicate with kn = 0: every output is proportional to kn
        contact_eval_nf<F32>(S.pd.grid, S.pose, valid ? S.pd.prm.kn : 0.0, xi1, sink);
    } else {
        sink.valid = valid;
        contact_eval<F32, DERIV, FRIC, DAMP>(S.pd.grid, S.pose, S.pd.prm, xi1, sink);
    }
}

__device__ __forceinline__ int warp_inclusive_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// position of the n-th (0-based) set bit of w; w has more than n set bits
__device__ __forceinline__ int nth_set_bit(uint32_t w, int n) {
    int pos = 0;
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        const int c = __popc((w >> pos) & ((1u << s) - 1u));
        if (n >= c) { n -= c; pos += s; }
    }
    return pos;
}

// The contacts of ranks [r0, r1) of the pair's mask row (set bits in word order), 32 at a time, starting in the
// detection tile whose first 64 mask words are `row` (wa / wb: this lane's two words of it, already loaded; base:
// the pair's contacts in front of the tile; nw: the pair's mask words from `row` on). Every lane finds ITS
// contact - rank c0 + lane - by a binary search over the scanned word counts (shuffles) and a bit select: no
// serial bit loop, no shared-memory id list. id0: point index of the first word of `row`.
template <bool F32, bool DERIV, bool FRIC, bool DAMP, bool NF, class Sink>
__device__ __forceinline__ void force_item(const ForceWarpShared<NF>& S, const uint32_t* __restrict__ row, int nw, int id0, uint32_t wa,
                                           uint32_t wb, int base, int r0, int r1, int lane, Sink& sink) {
    int c0 = r0, id = 0;
    for (int w = 0; w < nw; w += kDetectTileWords) {
        static_assert(kDetectTileWords == 64, "two words per lane and tile");
        if (w) {
            wa = (w + lane < nw) ? __ldg(row + w + lane) : 0u;
            wb = (w + 32 + lane < nw) ? __ldg(row + w + 32 + lane) : 0u;
        }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const uint32_t word = half ? wb : wa;
            const int cnt = __popc(word);
            const int incl = warp_inclusive_scan(cnt, lane);
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            const int cb = base;   // rank of the first contact of these 32 words
            base += tot;
            if (base <= c0) continue;
            for (;;) {
                const int wend = min(c0 + 32, r1);
                const int tgt = c0 + lane, rel = tgt - cb;
                const bool here = tgt < wend && rel >= 0 && rel < tot;
                const int relc = here ? rel : 0;
                int src = 0;   // first lane whose inclusive count exceeds relc
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const int v = __shfl_sync(0xffffffffu, incl, src + step - 1);
                    if (v <= relc) src += step;
                }
                const uint32_t ws = __shfl_sync(0xffffffffu, word, src);
                const int es = __shfl_sync(0xffffffffu, incl - cnt, src);
                if (here) id = id0 + (w + 32 * half + src) * 32 + nth_set_bit(ws, relc - es);
                if (base < wend) break;   // the round continues in the next 32 words
                const bool valid = tgt < wend;
                const int idf = __shfl_sync(0xffffffffu, id, 0);
                force_round<F32, DERIV, FRIC, DAMP, NF>(S, valid ? id : idf, valid, sink);
                c0 = wend;
                if (c0 >= r1) return;
            }
        }
    }
}

#ifndef SDFC_FORCE_MIN_BLOCKS
#define SDFC_FORCE_MIN_BLOCKS 2
#endif
#ifndef SDFC_FORCE_NF_MIN_BLOCKS
#define SDFC_FORCE_NF_MIN_BLOCKS 3  // planner fast path: 168 registers, 12 warps / SM
#endif
#ifndef SDFC_FORCE_VALUE_MIN_BLOCKS
#define SDFC_FORCE_VALUE_MIN_BLOCKS 4  // value-only kernel without friction / damping (line-search evaluations): 128 registers, 240 B of spills; 50.3 -> 47.6 ms per forward(100) of C2
#endif
template <bool DERIV, bool ANYFRIC, bool ANYDAMP>
constexpr int force_min_blocks() {
    return (ANYFRIC || ANYDAMP) ? SDFC_FORCE_MIN_BLOCKS : (DERIV ? SDFC_FORCE_NF_MIN_BLOCKS : SDFC_FORCE_VALUE_MIN_BLOCKS);
}

template <bool DERIV, bool ANYFRIC, bool ANYDAMP>
__global__ void __launch_bounds__(kForceThreads, force_min_blocks<DERIV, ANYFRIC, ANYDAMP>())
    k_force(BatchDev B, const double* __restrict__ E, const double* __restrict__ phi) {
    constexpr bool DM = ANYFRIC || ANYDAMP;
    constexpr bool NF = DERIV && !DM;  // planner fast path (contact_eval_nf)
    __shared__ ForceWarpShared<NF> SW[kNumWarps];
    const int lane = threadIdx.x & 31;
    ForceWarpShared<NF>& S = SW[threadIdx.x >> 5];
    const ForcePlanDev& P = B.plan[DERIV ? 1 : 0];
    const int n_items = P.hdr[0];
    const int total_warps = gridDim.x * kNumWarps;

    typename std::conditional<NF, ColSink, WarpSink<DERIV, DM>>::type sink;
    if constexpr (NF) sink.init(S.tile, S.acc, lane); else sink.init(lane);
    // Items are claimed from the list's counter by lane 0. The claim of the NEXT item is issued before this one is
    // evaluated and only read afterwards, so the atomic's round trip is off the critical path - except near the end
    // of the list, where holding an item in reserve would leave other warps idle.
    auto claim_raw = [&]() { return lane == 0 ? atomicAdd(P.hdr + 1, 1) : 0; };
    int cur_pair = -1;
    int item = __shfl_sync(0xffffffffu, claim_raw(), 0);
    while (item < n_items) {
        const int4 ta = __ldg(P.items + 2 * (size_t)item), tb = __ldg(P.items + 2 * (size_t)item + 1);
        const bool ahead = item + 2 * total_warps < n_items;
        int next_raw = 0;
        if (ahead) next_raw = claim_raw();
        const int ep = ta.x, r0 = ta.y, r1 = ta.z;
        const int pair = ep % B.n_pairs;
        // everything the item needs from global memory, requested at once: its tile's mask words, the two poses
        // (one element per lane), the pair record when it changes
        const uint32_t* row = B.mask + (unsigned)tb.x;
        const uint32_t wa = (lane < tb.y) ? __ldg(row + lane) : 0u;
        const uint32_t wb = (32 + lane < tb.y) ? __ldg(row + 32 + lane) : 0u;
        double pose_el = 0.0;
        if (lane < 24) {
            const int t = lane < 12 ? lane : lane - 12;   // element (r, c) of the top 3 x 4, column-major pose
            pose_el = __ldg(E + (size_t)(lane < 12 ? tb.z : tb.w) * 16 + (t % 3) + 4 * (t / 3));
        }
        double phi_el = 0.0;
        if constexpr (!NF) {
            if (lane < 12) phi_el = __ldg(phi + (size_t)(lane < 6 ? tb.z : tb.w) * 6 + (lane < 6 ? lane : lane - 6));
        }
        __syncwarp();   // the previous item's reads of the warp's workspace are done
        if (pair != cur_pair) {
            cur_pair = pair;
            const int* src = reinterpret_cast<const int*>(B.pairs + pair);
            int* dst = reinterpret_cast<int*>(&S.pd);
            for (int i = lane; i < (int)(sizeof(PairDev) / sizeof(int)); i += 32) dst[i] = __ldg(src + i);
        }
        if (lane < 24) {
            const int t = lane < 12 ? lane : lane - 12, r = t % 3, c = t / 3;
            double* R = lane < 12 ? S.pose.R1 : S.pose.R2;
            double* p = lane < 12 ? S.pose.p1 : S.pose.p2;
            if (c < 3) R[r * 3 + c] = pose_el; else p[r] = pose_el;
        }
        if constexpr (!NF) {
            if (lane < 6) S.pose.phi1[lane] = phi_el; else if (lane < 12) S.pose.phi2[lane - 6] = phi_el;
        }
        __syncwarp();
        if constexpr (NF) {
            if (lane < 9) {
                const int r = lane / 3, c = lane % 3;  // R21 = R2^T R1
                S.pose.R21[lane] = S.pose.R2[r] * S.pose.R1[c] + S.pose.R2[3 + r] * S.pose.R1[3 + c] + S.pose.R2[6 + r] * S.pose.R1[6 + c];
            }
            __syncwarp();
        }
        const int id0 = (int)((unsigned)tb.x - (unsigned)((size_t)(ep / B.n_pairs) * B.words_per_env + S.pd.word_offset)) * 32;
        const bool f32 = S.pd.grid.f32();
        const bool fric = ANYFRIC && (S.pd.prm.mu > kEps);        // .cpp:200,352
        const bool damp = ANYDAMP && (S.pd.prm.damping != 0.0);
        // warp-uniform variant dispatch (template flags prune what cannot occur)
#define SDFC_ITEM(F, D, FR, DA, N) force_item<F, D, FR, DA, N>(S, row, tb.y, id0, wa, wb, ta.w, r0, r1, lane, sink)
        if constexpr (NF) {
            if (f32) SDFC_ITEM(true, true, false, false, true); else SDFC_ITEM(false, true, false, false, true);
        } else if (f32) {
            if (fric) { if (damp) SDFC_ITEM(true, DERIV, ANYFRIC, ANYDAMP, false); else SDFC_ITEM(true, DERIV, ANYFRIC, false, false); }
            else      { if (damp) SDFC_ITEM(true, DERIV, false, ANYDAMP, false);  else SDFC_ITEM(true, DERIV, false, false, false); }
        } else {
            if (fric) { if (damp) SDFC_ITEM(false, DERIV, ANYFRIC, ANYDAMP, false); else SDFC_ITEM(false, DERIV, ANYFRIC, false, false); }
            else      { if (damp) SDFC_ITEM(false, DERIV, false, ANYDAMP, false);  else SDFC_ITEM(false, DERIV, false, false, false); }
        }
#undef SDFC_ITEM
        if constexpr (NF) flush_item(sink, S.pose.R21, P.partial + (size_t)item * P.stride, lane);
        else flush_item(sink, P.partial + (size_t)item * P.stride, lane);
        if (!ahead) next_raw = claim_raw();
        item = __shfl_sync(0xffffffffu, next_raw, 0);
    }
}

template <bool DERIV, bool ANYFRIC, bool ANYDAMP>
void launch_force_t(const BatchDev& B, const double* E, const double* phi, cudaStream_t stream) {
    // persistent: as many CTAs as stay resident, never more than the work list can feed
    const int resident = device_sm_count() * force_min_blocks<DERIV, ANYFRIC, ANYDAMP>();
    const int useful = (B.plan[DERIV ? 1 : 0].cap + kNumWarps - 1) / kNumWarps;
    k_force<DERIV, ANYFRIC, ANYDAMP><<<std::max(1, std::min(resident, useful)), kForceThreads, 0, stream>>>(B, E, phi);
}

}  // namespace sdfc
