// sdfc_kernels.cu — sm_100a kernels of the SDF-contact hot path.
//
//   k_detect_mask   LOOP C (collision_detection_general_SDF, CollisionDetection.cpp:124-147)
//                   for E environments x P pairs: one 1-bit/pt active mask word per
//                   warp-load of 32 points + an integer contact count per (env, pair).
//   k_force_plan    cuts the active contacts of all (env, pair)s into items of 32 G contacts
//                   (the work list of the force kernel, ForcePlanDev).
//   k_force         LOOP D (ForceGeneralSDFContact.cpp:147-447): persistent; a warp takes an
//                   item, picks its contacts out of the mask row, evaluates one contact per
//                   lane and reduces the 12 / 156 / 300 outputs across the warp with a register
//                   transpose-reduce; one partial per item.
//   k_reduce_items / k_scatter_system   fixed-order sums -> per-pair blocks and the
//                   per-environment system arrays (the += of .cpp:60-61,88-99).
//   k_query_points  per-point d / cell / active in original order (parity + sweep entry).
//   k_sdf_distance  raw trilinear sampling (SDFMesh::distance, Mesh.cpp:143-198).
//
// HBM-bound integer/fp64 work: no tensor cores. Points are read as coalesced SoA
// fp64; grid corners are gathered through L1/L2 (fp32-lossless copy when legal).
// Reductions are fixed-tree (no floating-point atomics): results are reproducible.
// Bit-critical arithmetic uses non-contractable intrinsics (see sdfc_device.cuh);
// the rest of the algebra is free to use FMAs.
#include "sdfc_kernels.cuh"

#include "sdfc_force.cuh"

namespace sdfc {
// the 8 variants are instantiated in sdfc_force_inst.cu (one TU each)
extern template void launch_force_t<false, false, false>(const BatchDev&, const double*, const double*, cudaStream_t);
extern template void launch_force_t<false, false, true>(const BatchDev&, const double*, const double*, cudaStream_t);
extern template void launch_force_t<false, true, false>(const BatchDev&, const double*, const double*, cudaStream_t);
extern template void launch_force_t<false, true, true>(const BatchDev&, const double*, const double*, cudaStream_t);
extern template void launch_force_t<true, false, false>(const BatchDev&, const double*, const double*, cudaStream_t);
extern template void launch_force_t<true, false, true>(const BatchDev&, const double*, const double*, cudaStream_t);
extern template void launch_force_t<true, true, false>(const BatchDev&, const double*, const double*, cudaStream_t);
extern template void launch_force_t<true, true, true>(const BatchDev&, const double*, const double*, cudaStream_t);
}  // namespace sdfc

namespace sdfc {

// --------------------------------------------------------------------------
// detection
// --------------------------------------------------------------------------
template <bool F32>
__device__ __forceinline__ double detect_point(const GridDev& G, const double* R1, const double* p1, const double* R2,
                                               const double* p2, double x0, double x1, double x2) {
    double xw[3], x[3];
    to_world(R1, p1, x0, x1, x2, xw);
    to_local(R2, p2, xw, x);
    return __dadd_rn(query_sdf<F32>(G, x[0], x[1], x[2], false, nullptr), G.col_th);  // BodySDFObj.cpp:212
}

// Skin state of this CTA's (env, pair): displacement of the relative pose since the reference.
struct SkinCta {
    bool rebuild;
    unsigned long long* stats;
    float out_margin;   // bmin + col_th rounded down, or -1 when that is not positive
    double dR[9], dp[3], fro, lip;
    float* margin;   // this pair's words of this environment
    int* fail;       // counters of (env, pair)
    int* elig;
    int* pair_margin;
};

// One thread per (env, pair): relative pose now, decide whether the margins are rebuilt (skin_decide_one, sdfc_kernels.cuh).
__global__ void k_skin_decide(BatchDev B, const double* __restrict__ E) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B.env_count * B.n_pairs) return;
    const int env = B.env0 + idx / B.n_pairs, pair = idx % B.n_pairs;
    if (B.env_mode != nullptr && B.env_mode[env] == 0) return;
    skin_decide_one(B, E, env, pair);
}

void launch_skin_decide(const BatchDev& B, const double* E, cudaStream_t stream) {
    const int n = B.env_count * B.n_pairs;
    k_skin_decide<<<(n + 127) / 128, 128, 0, stream>>>(B, E);
}

// Where the 8 corner values of a query come from. GlobalSampler: gathered through L1 / L2 (cell records,
// global bricks or the fp64 grid). BrickSampler (below): the tile's voxel brick staged in shared memory.
template <bool F32>
struct GlobalSampler {
    static constexpr bool kPrefetch = false;
    __device__ __forceinline__ double eval(const GridDev& G, const double* R1, const double* p1, const double* R2,
                                           const double* p2, double x0, double x1, double x2) {
        return detect_point<F32>(G, R1, p1, R2, p2, x0, x1, x2);
    }
};

template <bool F32, bool SKIN, class Sampler>
__device__ __forceinline__ void detect_tile(const PairDev& pd, const double* R1, const double* p1, const double* R2,
                                            const double* p2, int word0, uint32_t* mask_row, int* s_live, int* s_nlive,
                                            int* s_count, const SkinCta& sk, Sampler& smp, bool all_live = false, int word_end = -1) {
    if (word_end < 0) word_end = pd.n_words;   // first word past the range this call covers (a tile, or a part of one)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifndef SDFC_DETECT_U
#define SDFC_DETECT_U 1
#endif
    constexpr int U = SDFC_DETECT_U;  // words in flight per warp: all point loads of a group are issued before any use
    // Broad phase over the whole tile, one thread per word: transform the word's bounding
    // sphere centre into the SDF frame and test it against the grid box grown by the
    // radius. Rigid motions preserve distances, so a culled word has no point inside the
    // box (1e-7 absolute slack covers the rounding of the rotation); every such point
    // would evaluate to FLT_MAX + col_th > 0, i.e. inactive. Surviving words are compacted
    // (in word order) into shared memory and dealt round-robin to the warps, so that the
    // warps of a CTA finish together even when the contact region covers a few words only.
    if (all_live) {
        // the whole tile lies inside the grid box (CTA-uniform): every word is live, in word order
        const int nl = max(0, min(kDetectTileWords, word_end - word0));
        if (threadIdx.x < kDetectTileWords) s_live[kDetectTileWords + threadIdx.x] = word0 + threadIdx.x;
        if (threadIdx.x == 0) { s_nlive[0] = min(nl, 32); s_nlive[1] = max(nl - 32, 0); }
    } else if (threadIdx.x < kDetectTileWords) {
        const int w = word0 + threadIdx.x;
        bool keep = false, culled = false, used_up = false, elig_out = false;
        if (w < word_end) {
            const double* sp = pd.wsphere + 4 * (size_t)w;
            double cw[3], cl[3];
            to_world(R1, p1, __ldg(sp), __ldg(sp + 1), __ldg(sp + 2), cw);
            to_local(R2, p2, cw, cl);
            const double r = __ldg(sp + 3) + 1e-7;
            keep = cl[0] >= pd.grid.mn[0] - r && cl[0] <= pd.grid.mx[0] + r &&
                   cl[1] >= pd.grid.mn[1] - r && cl[1] <= pd.grid.mx[1] + r &&
                   cl[2] >= pd.grid.mn[2] - r && cl[2] <= pd.grid.mx[2] + r;
            if constexpr (SKIN) {
                if (sk.rebuild) {
                    // every point outside the box at the reference pose: they all carry the outside margin
                    if (!keep) { sk.margin[w] = sk.out_margin; elig_out = sk.out_margin >= 0.f; }
                } else if (keep) {
                    const float m = sk.margin[w];
                    if (m >= 0.f) {
                        // displacement bound of every point of the word since the reference pose:
                        // centre moved by |dR c + dp|, the rest of the sphere by at most |dR|_F r
                        const double cx = __ldg(sp), cy = __ldg(sp + 1), cz = __ldg(sp + 2);
                        const double ux = sk.dR[0] * cx + sk.dR[1] * cy + sk.dR[2] * cz + sk.dp[0];
                        const double uy = sk.dR[3] * cx + sk.dR[4] * cy + sk.dR[5] * cz + sk.dp[1];
                        const double uz = sk.dR[6] * cx + sk.dR[7] * cy + sk.dR[8] * cz + sk.dp[2];
                        const double delta = sqrt(ux * ux + uy * uy + uz * uz) + sk.fro * __ldg(sp + 3);
                        if ((double)m - sk.lip * delta * (1.0 + 1e-9) > 1e-9) { keep = false; culled = true; }  // provably no contact
                        else used_up = true;
                    }
                }
            }
            if (!keep) mask_row[pd.word_offset + w] = 0u;  // culled words: empty mask
        }
        if (SKIN && sk.rebuild) {
            const unsigned eb = __ballot_sync(0xffffffffu, elig_out);
            const unsigned ob = __ballot_sync(0xffffffffu, w < word_end && !keep);
            if (lane == 0 && eb) atomicAdd(sk.elig, __popc(eb));
            // signed-int order == float order on {-1} u [0, inf): one word without a margin makes the pair's -1
            if (lane == 0 && ob) atomicMin(sk.pair_margin, __float_as_int(sk.out_margin));
        }
        if (SKIN && !sk.rebuild) {   // warp-uniform
            const unsigned cb = __ballot_sync(0xffffffffu, culled), ub = __ballot_sync(0xffffffffu, used_up);
            if (lane == 0) {
                if (cb) atomicAdd(sk.stats, (unsigned long long)__popc(cb));
                if (ub) { atomicAdd(sk.fail, __popc(ub)); atomicAdd(sk.stats + 1, (unsigned long long)__popc(ub)); }
            }
        }
        const unsigned live = __ballot_sync(0xffffffffu, keep);
        // kDetectTileWords == 64: warps 0 and 1 hold the two ballots
        if (lane == 0) s_nlive[warp] = __popc(live);
        __syncwarp();
        s_live[32 * warp + lane] = keep ? (__popc(live & ((1u << lane) - 1u))) : -1;  // rank inside this warp's ballot
    }
    __syncthreads();
    const int n0 = s_nlive[0], nlive = n0 + s_nlive[1];
    // rank -> word: scatter into a dense list (second half of s_live)
    int* list = s_live + kDetectTileWords;
    if (!all_live) {
        if (threadIdx.x < kDetectTileWords) {
            const int rk = s_live[threadIdx.x];
            if (rk >= 0) list[rk + (threadIdx.x >= 32 ? n0 : 0)] = word0 + threadIdx.x;
        }
        __syncthreads();
    }
    int cnt = 0;
    // software pipeline: the coordinates of the warp's next group are requested before the current group is
    // evaluated, so a warp never sits on an exposed global load between two words
    const int kStep = (int)(blockDim.x >> 5) * U;   // warps of this CTA x words in flight per warp
    int nws[U];
    double npx[U], npy[U], npz[U];
    bool ninr[U];
    auto fetch = [&](int k) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            nws[u] = (k + u < nlive) ? list[k + u] : -1;
            const int idx = nws[u] * 32 + lane;
            ninr[u] = (nws[u] >= 0) && (idx < pd.n);
            npx[u] = ninr[u] ? __ldg(pd.sx + idx) : 0.0;
            npy[u] = ninr[u] ? __ldg(pd.sy + idx) : 0.0;
            npz[u] = ninr[u] ? __ldg(pd.sz + idx) : 0.0;
        }
    };
    if (Sampler::kPrefetch) fetch(warp * U);
    for (int k = warp * U; k < nlive; k += kStep) {
        int ws[U];
        double px[U], py[U], pz[U];
        bool inr[U];
        if (!Sampler::kPrefetch) fetch(k);   // cache-resident grids: 40 registers / 6 CTAs per SM matter more than the overlap
#pragma unroll
        for (int u = 0; u < U; ++u) { ws[u] = nws[u]; px[u] = npx[u]; py[u] = npy[u]; pz[u] = npz[u]; inr[u] = ninr[u]; }
        if (Sampler::kPrefetch && k + kStep < nlive) fetch(k + kStep);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ws[u] < 0) continue;  // warp-uniform
            double d = 1e300;
            if (inr[u]) d = smp.eval(pd.grid, R1, p1, R2, p2, px[u], py[u], pz[u]);
            const bool act = inr[u] && d < 0.;
            const unsigned m = __ballot_sync(0xffffffffu, act);
            if (lane == 0) { mask_row[pd.word_offset + ws[u]] = m; cnt += __popc(m); }
            if (SKIN && sk.rebuild) {
                // margin of the word at the reference pose: smallest d of its points; a point outside the
                // box (d = FLT_MAX + col_th) carries the outside margin: before it can be active it has to
                // cross a face of the box, where the interpolant is >= GridDev::bmin
                double dm = inr[u] ? (d > 1e37 ? (double)sk.out_margin : d) : 1e300;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dm = fmin(dm, __shfl_xor_sync(0xffffffffu, dm, o));
                if (lane == 0) {
                    const bool ok = dm > 0.0;
                    const float mw = ok ? __double2float_rd(dm) : -1.f;
                    sk.margin[ws[u]] = mw;
                    atomicMin(sk.pair_margin, __float_as_int(mw));
                    if (ok) atomicAdd(sk.elig, 1);
                }
            }
        }
    }
    if (lane == 0 && cnt) atomicAdd(s_count, cnt);
}

#ifndef SDFC_DETECT_MIN_BLOCKS
#define SDFC_DETECT_MIN_BLOCKS 6  // 40 registers: the kernel is latency-bound on the cell gathers, occupancy wins
#endif
// SKIN = false is the plain kernel (every batch evaluation outside the stepper); SKIN = true adds the
// stepper's margin test / margin rebuild, compiled out of the plain instantiation.
template <bool SKIN>
__global__ void __launch_bounds__(SKIN ? kDetectThreadsSkin : kDetectThreadsPlain,
                                  SDFC_DETECT_MIN_BLOCKS * kDetectThreads / (SKIN ? kDetectThreadsSkin : kDetectThreadsPlain))
    k_detect_mask(BatchDev B, const double* __restrict__ E) {
    __shared__ PairDev pd;
    __shared__ double R1[9], p1[3], R2[9], p2[3];
    __shared__ int s_count;
    __shared__ int s_live[2 * kDetectTileWords];
    __shared__ int s_nlive[kDetectThreads / 32];
    static_assert(kDetectTileWords == 64, "the broad phase compaction assumes two ballots per tile");
    const int tile = blockIdx.x, env = blockIdx.y + B.env0;
    if (B.env_mode != nullptr && __ldg(B.env_mode + env) == 0) return;  // CTA-uniform
    const int pair = __ldg(B.tile_pair + tile), word0 = __ldg(B.tile_word0 + tile);
    // whole pair provably contact-free (k_skin_decide): n_active stays 0, so nothing reads this pair's mask words
    if (SKIN && B.skin_mode[(size_t)env * B.n_pairs + pair] == 2) return;  // CTA-uniform
    {
        const int* src = reinterpret_cast<const int*>(B.pairs + pair);
        int* dst = reinterpret_cast<int*>(&pd);
        for (int i = threadIdx.x; i < (int)(sizeof(PairDev) / sizeof(int)); i += blockDim.x) dst[i] = __ldg(src + i);
        if (threadIdx.x == 0) s_count = 0;
    }
    __syncthreads();
    __shared__ SkinCta sk;
    if (threadIdx.x < 12) stage_pose(E, B.n_bodies, env, pd.body_a, R1, p1, threadIdx.x);
    else if (threadIdx.x >= 32 && threadIdx.x < 44) stage_pose(E, B.n_bodies, env, pd.body_b, R2, p2, threadIdx.x - 32);
    else if (SKIN && threadIdx.x == 64) {
        {
            const size_t ep = (size_t)env * B.n_pairs + pair;
            sk.rebuild = B.skin_mode[ep] != 0;
            double f2 = 0.0;
            for (int i = 0; i < 12; ++i) {
                const double dlt = B.skin_cur[ep * 12 + i] - B.skin_ref[ep * 12 + i];
                if (i < 9) { sk.dR[i] = dlt; f2 += dlt * dlt; } else sk.dp[i - 9] = dlt;
            }
            sk.fro = sqrt(f2);
            sk.lip = pd.grid.lip;
            const double om = pd.grid.bmin + pd.grid.col_th;
            sk.out_margin = om > 0.0 ? __double2float_rd(fmin(om, 1e30)) : -1.f;
            sk.margin = B.skin_margin + (size_t)env * B.words_per_env + pd.word_offset;
            sk.stats = B.skin_stats;
            sk.fail = B.skin_fail + ep;
            sk.elig = B.skin_elig + ep;
            sk.pair_margin = B.skin_pair_margin + ep;
        }
    }
    __syncthreads();
    uint32_t* mask_row = B.mask + (size_t)env * B.words_per_env;
    if (pd.grid.f32()) { GlobalSampler<true> smp; detect_tile<true, SKIN>(pd, R1, p1, R2, p2, word0, mask_row, s_live, s_nlive, &s_count, sk, smp); }
    else { GlobalSampler<false> smp; detect_tile<false, SKIN>(pd, R1, p1, R2, p2, word0, mask_row, s_live, s_nlive, &s_count, sk, smp); }
    __syncthreads();
    if (threadIdx.x == 0) {
        B.tile_count[(size_t)env * B.n_tiles + tile] = s_count;   // the force kernel finds an item's contacts tile by tile
        if (s_count) {
            atomicAdd(B.n_active + (size_t)env * B.n_pairs + pair, s_count);
            atomicAdd(B.plan[0].hdr + 2, (s_count + 31) >> 5);   // bound of the force kernels' 32-contact rounds
        }
    }
}

// --------------------------------------------------------------------------
// detection with the tile's voxel brick staged in shared memory by TMA
// --------------------------------------------------------------------------
// For grids that do not stay in L2 (GridDev::b32, 4 B per node) a gather per query costs a full DRAM
// line per corner group. Instead the CTA of a tile (2048 spatially sorted points) maps the tile's box
// (PairDev::tile_box, contact-body frame) into the SDF frame, takes the bricks under it, and pulls them
// into shared memory with cp.async.bulk.tensor (one 4-D box of 64 x nbx x nby x 1 per brick layer,
// completion on an mbarrier) while the broad phase and the point transforms run. Every query then reads
// its 8 corners from shared memory: same values, same lerp sequence, same bits as BodySDFObj.cpp:132-153.
// A tile whose brick does not fit (or a pair without bricks) gathers from global memory as before.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_bricks(const void* tmap, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

struct BrickSampler {
    static constexpr bool kPrefetch = true;
    uint32_t sb;         // shared-memory address of the staged bricks [nbz][nby][nbx][64], 0 = not staged (global gathers)
    uint64_t* bar;
    int bx0, by0, bz0, nbx, nby, nbz;
    bool waited;
    bool all_inside;     // the image of the tile's box lies inside [mn, mx]: every point samples the grid
    uint32_t parity;     // phase of the mbarrier this stage completes
    __device__ __forceinline__ double eval(const GridDev& G, const double* R1, const double* p1, const double* R2,
                                           const double* p2, double x0, double x1, double x2) {
        double xw[3], x[3];
        to_world(R1, p1, x0, x1, x2, xw);
        to_local(R2, p2, xw, x);
        if (!all_inside && !in_box(G, x[0], x[1], x[2])) return __dadd_rn((double)FLT_MAX, G.col_th);   // BodySDFObj.cpp:170,212
        double ci = exact_div(__dsub_rn(x[0], G.mn[0]), G.dx, G.rdx);
        double cj = exact_div(__dsub_rn(x[1], G.mn[1]), G.dy, G.rdy);
        double ck = exact_div(__dsub_rn(x[2], G.mn[2]), G.dz, G.rdz);
        int i = (int)floor(ci), j = (int)floor(cj), k = (int)floor(ck);
        ci = __dsub_rn(ci, (double)i); cj = __dsub_rn(cj, (double)j); ck = __dsub_rn(ck, (double)k);
        i = max(0, min(i, G.ni - 1)); j = max(0, min(j, G.nj - 1)); k = max(0, min(k, G.nk - 1));   // memory guard, as load_cell
        double v[8];
        const int i1 = i + 1, j1 = j + 1, k1 = k + 1;   // the padded layer replicates the clamp of BodySDFObj.cpp:137-140
        const int rx0 = (i >> 2) - bx0, rx1 = (i1 >> 2) - bx0, ry0 = (j >> 2) - by0, ry1 = (j1 >> 2) - by0;
        const int rz0 = (k >> 2) - bz0, rz1 = (k1 >> 2) - bz0;
        // the brick covers the image of the tile's box, so this always holds for a staged tile; checked anyway
        const bool staged = sb != 0u && rx0 >= 0 && rx1 < nbx && ry0 >= 0 && ry1 < nby && rz0 >= 0 && rz1 < nbz;
        if (staged) {
            if (!waited) { mbar_wait(bar, parity); waited = true; }
            const int sy = nbx * 64, sz = sy * nby;
            const int ox0 = rx0 * 64 + (i & 3), ox1 = rx1 * 64 + (i1 & 3);
            const int oy0 = ry0 * sy + (j & 3) * 4, oy1 = ry1 * sy + (j1 & 3) * 4;
            const int oz0 = rz0 * sz + (k & 3) * 16, oz1 = rz1 * sz + (k1 & 3) * 16;
            const uint32_t a00 = sb + 4u * (uint32_t)(oy0 + oz0), a10 = sb + 4u * (uint32_t)(oy1 + oz0);
            const uint32_t a01 = sb + 4u * (uint32_t)(oy0 + oz1), a11 = sb + 4u * (uint32_t)(oy1 + oz1);
            const uint32_t x0b = 4u * (uint32_t)ox0, x1b = 4u * (uint32_t)ox1;
            v[0] = (double)lds_f32(a00 + x0b); v[1] = (double)lds_f32(a00 + x1b);
            v[2] = (double)lds_f32(a10 + x0b); v[3] = (double)lds_f32(a10 + x1b);
            v[4] = (double)lds_f32(a01 + x0b); v[5] = (double)lds_f32(a01 + x1b);
            v[6] = (double)lds_f32(a11 + x0b); v[7] = (double)lds_f32(a11 + x1b);
        } else {
            load_cell<true>(G, i, j, k, v);
        }
        return __dadd_rn(lerp8(v, ci, cj, ck), G.col_th);
    }
};

#ifndef SDFC_BRICK_MIN_BLOCKS
#define SDFC_BRICK_MIN_BLOCKS 3
#endif
#ifndef SDFC_BRICK_THREADS
#define SDFC_BRICK_THREADS 256
#endif
constexpr int kBrickThreads = SDFC_BRICK_THREADS;
__global__ void __launch_bounds__(kBrickThreads, SDFC_BRICK_MIN_BLOCKS) k_detect_brick(BatchDev B, const double* __restrict__ E) {
    extern __shared__ __align__(128) unsigned char brick_raw[];
    __shared__ PairDev pd;
    __shared__ double R1[9], p1[3], R2[9], p2[3];
    __shared__ int s_count;
    __shared__ int s_live[2 * kDetectTileWords];
    __shared__ int s_nlive[kBrickThreads / 32];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_brick[8];   // bx0, by0, bz0, nbx, nby, nbz, staged, tile image inside the grid box
    const int tile = blockIdx.x, env = blockIdx.y + B.env0;
    if (B.env_mode != nullptr && __ldg(B.env_mode + env) == 0) return;
    const int pair = __ldg(B.tile_pair + tile), word0 = __ldg(B.tile_word0 + tile);
    {
        const int* src = reinterpret_cast<const int*>(B.pairs + pair);
        int* dst = reinterpret_cast<int*>(&pd);
        for (int i = threadIdx.x; i < (int)(sizeof(PairDev) / sizeof(int)); i += blockDim.x) dst[i] = __ldg(src + i);
        if (threadIdx.x == 0) { s_count = 0; s_brick[6] = 0; s_brick[7] = 0; mbar_init(&s_bar, 1); }
    }
    __syncthreads();
    if (threadIdx.x < 12) stage_pose(E, B.n_bodies, env, pd.body_a, R1, p1, threadIdx.x);
    else if (threadIdx.x >= 32 && threadIdx.x < 44) stage_pose(E, B.n_bodies, env, pd.body_b, R2, p2, threadIdx.x - 32);
    __syncthreads();
    float* sb = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(brick_raw) + 127) & ~(uintptr_t)127);
    const bool bricks = pd.grid.b32 != nullptr && pd.grid.tmaps != nullptr && pd.tile_box != nullptr;
    if (bricks && threadIdx.x >= 32 && threadIdx.x < 35) {
        // the tile's coordinates (3 contiguous runs of the SoA copy) are pulled into L2 now, so that the warps'
        // loads below find them there instead of waiting on DRAM one word at a time
        const double* base = (threadIdx.x == 32) ? pd.sx : (threadIdx.x == 33 ? pd.sy : pd.sz);
        const int i0 = word0 * 32, cnt = min(pd.n - i0, kDetectTileWords * 32);
        const uint32_t bytes = (uint32_t)(cnt * 8) & ~15u;
        if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + i0), "r"(bytes) : "memory");
    }
    uint32_t* mask_row = B.mask + (size_t)env * B.words_per_env;
    SkinCta sk_unused;
    if (!bricks) {
        if (pd.grid.f32()) {
            GlobalSampler<true> smp;
            detect_tile<true, false>(pd, R1, p1, R2, p2, word0, mask_row, s_live, s_nlive, &s_count, sk_unused, smp);
        } else {
            GlobalSampler<false> smp;
            detect_tile<false, false>(pd, R1, p1, R2, p2, word0, mask_row, s_live, s_nlive, &s_count, sk_unused, smp);
        }
    } else {
        // The tile, or - when the bricks under its box exceed the shared-memory cap (sparse point sets: 2048 points
        // spread over many cells) - its two halves, or their quarters, one after the other: part `s` of the tile
        // (PairDev::tile_box holds 7 boxes per tile: 0 = all 64 words, 1..2 = halves, 3..6 = quarters) stages the
        // bricks under ITS box and is evaluated from shared memory; only a quarter that still does not fit gathers
        // from global memory. One mbarrier, one phase per staged part.
        uint32_t stage = 0;
        // returns false when part s has to be split (nothing evaluated)
        auto run_part = [&](int s, bool leaf) -> bool {
            const int len = s == 0 ? kDetectTileWords : (s < 3 ? kDetectTileWords / 2 : kDetectTileWords / 4);
            const int w0 = word0 + (s == 0 ? 0 : (s < 3 ? (s - 1) * len : (s - 3) * len));
            const int w1 = min(pd.n_words, w0 + len);
            if (w0 >= pd.n_words) return true;   // CTA-uniform: the part lies past the last word
            if (threadIdx.x < 32) {
                // image of the part's box in the SDF frame: min / max over its 8 corners (lanes 0..7)
                const double* tb = pd.tile_box + 6 * ((size_t)(word0 / kDetectTileWords) * 7 + s);
                const int c = threadIdx.x & 7;
                double cw[3], cl[3];
                to_world(R1, p1, __ldg(tb + ((c & 1) ? 3 : 0)), __ldg(tb + ((c & 2) ? 4 : 1)), __ldg(tb + ((c & 4) ? 5 : 2)), cw);
                to_local(R2, p2, cw, cl);
                double lo[3] = {cl[0], cl[1], cl[2]}, hi[3] = {cl[0], cl[1], cl[2]};
#pragma unroll
                for (int o = 4; o > 0; o >>= 1)
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
                        hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
                    }
                if (threadIdx.x == 0) {
                    // only points inside [mn, mx] sample the grid; 1e-7 covers the rounding of the transforms (as the broad phase)
                    const GridDev& G = pd.grid;
                    const double dd[3] = {G.dx, G.dy, G.dz}, rd[3] = {G.rdx, G.rdy, G.rdz};
                    const int nn[3] = {G.ni, G.nj, G.nk}, nbr[3] = {G.nbx, G.nby, G.nbz};
                    int b0[3], nb[3];
                    bool any = true, inside = true;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double l = fmax(lo[a] - 1e-7, G.mn[a]), h = fmin(hi[a] + 1e-7, G.mx[a]);
                        any = any && (l <= h);
                        inside = inside && (lo[a] - 1e-7 >= G.mn[a]) && (hi[a] + 1e-7 <= G.mx[a]);
                        // cell index is monotonic in the coordinate (correctly rounded quotient, then floor)
                        int c0 = (int)floor(exact_div(__dsub_rn(l, G.mn[a]), dd[a], rd[a]));
                        int c1 = (int)floor(exact_div(__dsub_rn(h, G.mn[a]), dd[a], rd[a]));
                        c0 = max(0, min(c0, nn[a] - 1)); c1 = max(0, min(c1, nn[a] - 1));
                        b0[a] = c0 >> 2;
                        nb[a] = min(((c1 + 1) >> 2), nbr[a] - 1) - b0[a] + 1;   // + 1: upper neighbour node
                    }
                    const bool fits = any && nb[0] <= kBrickMaxXY && nb[1] <= kBrickMaxXY &&
                                      (size_t)nb[0] * nb[1] * nb[2] * 256 <= (size_t)B.brick_cap;
                    const bool split = any && !fits && !leaf;
                    if (B.brick_stats != nullptr && !split) atomicAdd(B.brick_stats + (fits ? 0 : (any ? 1 : 2)), 1ull);
                    s_brick[0] = b0[0]; s_brick[1] = b0[1]; s_brick[2] = b0[2];
                    s_brick[3] = nb[0]; s_brick[4] = nb[1]; s_brick[5] = nb[2];
                    s_brick[6] = fits ? 1 : (split ? 2 : 0);
                    s_brick[7] = (any && inside) ? 1 : 0;
                    if (fits) {
                        const uint32_t slice = (uint32_t)nb[0] * nb[1] * 256;
                        mbar_arrive_expect_tx(&s_bar, slice * nb[2]);
                        const unsigned char* tm = static_cast<const unsigned char*>(G.tmaps) +
                                                  (size_t)((nb[0] - 1) * kBrickMaxXY + (nb[1] - 1)) * 128;
                        for (int z = 0; z < nb[2]; ++z)
                            tma_load_bricks(tm, &s_bar, reinterpret_cast<unsigned char*>(sb) + (size_t)z * slice, 0, b0[0], b0[1], b0[2] + z);
                    }
                }
            }
            __syncthreads();
            const int how = s_brick[6];
            if (how == 2) { __syncthreads(); return false; }   // every thread has read the verdict before it is rewritten
            const bool staged = how == 1;
            BrickSampler smp{staged ? smem_u32(sb) : 0u, &s_bar, s_brick[0], s_brick[1], s_brick[2], s_brick[3], s_brick[4], s_brick[5], false,
                             s_brick[7] != 0, stage & 1u};
            detect_tile<true, false>(pd, R1, p1, R2, p2, w0, mask_row, s_live, s_nlive, &s_count, sk_unused, smp, s_brick[7] != 0, w1);
            if (staged) {
                if (!smp.waited) mbar_wait(&s_bar, stage & 1u);   // no part (and not the CTA) ends under an in-flight bulk copy
                stage += 1;
            }
            __syncthreads();   // bricks, live-word lists and the verdict are consumed
            return true;
        };
        if (!run_part(0, false))
            for (int h = 0; h < 2; ++h)
                if (!run_part(1 + h, false))
                    for (int q = 0; q < 2; ++q) run_part(3 + 2 * h + q, true);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        B.tile_count[(size_t)env * B.n_tiles + tile] = s_count;   // the force kernel finds an item's contacts tile by tile
        if (s_count) {
            atomicAdd(B.n_active + (size_t)env * B.n_pairs + pair, s_count);
            atomicAdd(B.plan[0].hdr + 2, (s_count + 31) >> 5);   // bound of the force kernels' 32-contact rounds
        }
    }
}

void launch_detect_mask(const BatchDev& B, const double* E, cudaStream_t stream, bool reset_counts) {
    if (reset_counts) {
        cudaMemsetAsync(B.n_active + (size_t)B.env0 * B.n_pairs, 0, sizeof(int32_t) * (size_t)B.env_count * B.n_pairs, stream);
        cudaMemsetAsync(B.plan[0].hdr, 0, sizeof(int) * 3, stream);   // work lists of the force kernels: empty
        cudaMemsetAsync(B.plan[1].hdr, 0, sizeof(int) * 3, stream);
    }
    if (B.n_tiles == 0 || B.env_count == 0) return;   // every pair has an empty point set: all counts stay 0
    dim3 grid(B.n_tiles, B.env_count);
    if (B.any_bricks && B.skin_margin == nullptr) {
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(k_detect_brick, cudaFuncAttributeMaxDynamicSharedMemorySize, kBrickSmemMax + 128);
            configured = true;
        }
        k_detect_brick<<<grid, kBrickThreads, B.brick_cap + 128, stream>>>(B, E);
        return;
    }
    if (B.skin_margin != nullptr) k_detect_mask<true><<<grid, kDetectThreadsSkin, 0, stream>>>(B, E);
    else k_detect_mask<false><<<grid, kDetectThreadsPlain, 0, stream>>>(B, E);
}

// --------------------------------------------------------------------------
// work list of the force kernels
// --------------------------------------------------------------------------
constexpr int kPlanWarps = 8;

// One warp per (environment, pair) with contacts. G = the smallest power-of-two multiple of g_min for which the
// list is certain to fit (items <= rounds / G + pairs; the number of rounds is bounded by the sum detection
// accumulated tile by tile, so every warp derives the same G without a grid-wide pass). The pair's items take
// consecutive slots handed out by one atomic - their position in the list varies from run to run, the sums do
// not: a pair's partials are added in item order (reduce_items_one). Lane k fills item k, k + 32, ...: its
// contact range, the detection tile that holds its first contact (binary search over the scanned tile counts)
// and the addresses the force kernel starts from.
__global__ void __launch_bounds__(kPlanWarps * 32) k_force_plan(BatchDev B, int deriv_sel) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * kPlanWarps + (threadIdx.x >> 5);
    if (i >= B.env_count * B.n_pairs) return;
    const int env = B.env0 + i / B.n_pairs, pair = i % B.n_pairs;
    int sel = deriv_sel;
    if (B.env_mode != nullptr) {
        const int m = __ldg(B.env_mode + env);
        if (m == 0) return;
        sel = (m == 2) ? 1 : 0;
    }
    const int ep = env * B.n_pairs + pair;
    const int n_act = B.n_active[ep];
    if (n_act == 0) return;
    const ForcePlanDev& P = B.plan[sel];
    int G = P.g_min;
    {
        const long long bound = B.plan[0].hdr[2];
        const int n_ep = B.env_count * B.n_pairs;
        while ((bound + G - 1) / G + n_ep > P.cap) G *= 2;
    }
    const int per_item = 32 * G;
    const int n_items = (n_act + per_item - 1) / per_item;
    int item0 = 0;
    if (lane == 0) {
        item0 = atomicAdd(P.hdr, n_items);
        P.pair_item0[ep] = item0;
        P.hdr[3] = G;
    }
    item0 = __shfl_sync(0xffffffffu, item0, 0);
    const PairDev& pd = B.pairs[pair];
    const int32_t* tc = B.tile_count + (size_t)env * B.n_tiles + pd.tile_offset;
    const int nt = (pd.n_words + kDetectTileWords - 1) / kDetectTileWords;
    const unsigned row0 = (unsigned)((size_t)env * B.words_per_env + pd.word_offset);
    const int slot_a = env * B.n_bodies + pd.body_a, slot_b = env * B.n_bodies + pd.body_b;
    int k0 = 0, base = 0;   // items below k0 are written; contacts in front of tile chunk t0
    for (int t0 = 0; t0 < nt && k0 < n_items; t0 += 32) {
        const int c = (t0 + lane < nt) ? tc[t0 + lane] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        // items whose first contact lies in these 32 tiles: k0 <= k < k1
        const int k1 = min(n_items, (base + tot + per_item - 1) / per_item);
        for (int kb = k0; kb < k1; kb += 32) {
            const int k = kb + lane;
            const bool mine = k < k1;
            const int rel = mine ? k * per_item - base : 0;   // 0 <= rel < tot
            int src = 0;   // first tile of the chunk whose inclusive count exceeds rel
#pragma unroll
            for (int step = 16; step; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, incl, src + step - 1);
                if (v <= rel) src += step;
            }
            const int before = __shfl_sync(0xffffffffu, incl - c, src);
            if (mine) {
                const int r0 = k * per_item, t = t0 + src;
                P.items[2 * (size_t)(item0 + k)] = make_int4(ep, r0, min(n_act, r0 + per_item), base + before);
                P.items[2 * (size_t)(item0 + k) + 1] =
                    make_int4((int)(row0 + (unsigned)t * kDetectTileWords), pd.n_words - t * kDetectTileWords, slot_a, slot_b);
            }
        }
        k0 = k1;
        base += tot;
    }
}

void launch_force_plan(const BatchDev& B, bool deriv, cudaStream_t stream) {
    const int n_ep = B.env_count * B.n_pairs;
    if (n_ep == 0) return;
    k_force_plan<<<(n_ep + kPlanWarps - 1) / kPlanWarps, kPlanWarps * 32, 0, stream>>>(B, deriv ? 1 : 0);
}

int device_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

void launch_force(const BatchDev& B, const double* E, const double* phi, bool deriv, bool any_fric, bool any_damp,
                  cudaStream_t stream) {
    if (deriv) {
        if (any_fric) { if (any_damp) launch_force_t<true, true, true>(B, E, phi, stream); else launch_force_t<true, true, false>(B, E, phi, stream); }
        else          { if (any_damp) launch_force_t<true, false, true>(B, E, phi, stream); else launch_force_t<true, false, false>(B, E, phi, stream); }
    } else {
        if (any_fric) { if (any_damp) launch_force_t<false, true, true>(B, E, phi, stream); else launch_force_t<false, true, false>(B, E, phi, stream); }
        else          { if (any_damp) launch_force_t<false, false, true>(B, E, phi, stream); else launch_force_t<false, false, false>(B, E, phi, stream); }
    }
}

// --------------------------------------------------------------------------
// finalize: item partials -> per-pair blocks -> system arrays
// --------------------------------------------------------------------------
// Per-pair block layout (natural): fm_sub[12], Km_sub[144] col-major, Dm_sub[144] col-major.
__global__ void k_reduce_items(BatchDev B, int deriv, int dm, double* __restrict__ fm_sub, double* __restrict__ Km_sub,
                                double* __restrict__ Dm_sub) {
    const int pair = blockIdx.x, env = blockIdx.y + B.env0;
    if (B.env_mode != nullptr) {
        const int m = __ldg(B.env_mode + env);
        if (m == 0) return;
        deriv = (m == 2) ? 1 : 0;
    }
    reduce_items_one(B, env, pair, deriv, dm, threadIdx.x, blockDim.x, fm_sub, Km_sub, Dm_sub);
}

// Gather form of the reference's scatter (.cpp:60-61,88-99): one thread per system
// entry walks the pairs in force order, so the accumulation order is the reference's.
__device__ __forceinline__ void scatter_env(const BatchDev& B, int env, int deriv, const double* fm_sub, const double* Km_sub,
                                            const double* Dm_sub, double* fm, double* Km, double* Dm, int tid, int nthreads) {
    const int nm = 6 * B.n_bodies;
    const size_t e0 = (size_t)env * B.n_pairs;
    for (int i = tid; i < nm; i += nthreads) {
        const int body = i / 6, r = i % 6;
        double s = 0.0;
        for (int p = 0; p < B.n_pairs; ++p) {
            const int a = B.pairs[p].body_a, b = B.pairs[p].body_b;
            if (a == body) s += fm_sub[(e0 + p) * 12 + r];
            if (b == body) s += fm_sub[(e0 + p) * 12 + 6 + r];
        }
        fm[(size_t)env * nm + i] = s;
    }
    if (!deriv) return;
    for (int i = tid; i < nm * nm; i += nthreads) {
        const int R = i % nm, C = i / nm;  // column-major
        const int bR = R / 6, r = R % 6, bC = C / 6, c = C % 6;
        double sk = 0.0, sd = 0.0;
        for (int p = 0; p < B.n_pairs; ++p) {
            const int bodies[2] = {B.pairs[p].body_a, B.pairs[p].body_b};
#pragma unroll
            for (int blk = 0; blk < 4; ++blk) {  // topLeft, topRight, bottomLeft, bottomRight
                if (bodies[blk >> 1] == bR && bodies[blk & 1] == bC) {
                    const size_t src = (e0 + p) * 144 + (r + 6 * (blk >> 1)) + 12 * (c + 6 * (blk & 1));
                    sk += Km_sub[src];
                    if (Dm != nullptr) sd += Dm_sub[src];
                }
            }
        }
        Km[(size_t)env * nm * nm + i] = sk;
        if (Dm != nullptr) Dm[(size_t)env * nm * nm + i] = sd;
    }
}

__global__ void k_scatter_system(BatchDev B, int deriv, const double* __restrict__ fm_sub,
                                 const double* __restrict__ Km_sub, const double* __restrict__ Dm_sub,
                                 double* __restrict__ fm, double* __restrict__ Km, double* __restrict__ Dm) {
    scatter_env(B, blockIdx.x + B.env0, deriv, fm_sub, Km_sub, Dm_sub, fm, Km, Dm, threadIdx.x, blockDim.x);
}

// Both steps in one launch, a CTA per environment: the pair blocks of the environment, then its system arrays.
__global__ void __launch_bounds__(256) k_reduce_scatter(BatchDev B, int deriv, int dm, double* fm_sub, double* Km_sub, double* Dm_sub,
                                                        double* fm, double* Km, double* Dm) {
    const int env = blockIdx.x + B.env0;
    for (int p = 0; p < B.n_pairs; ++p) reduce_items_one(B, env, p, deriv, dm, threadIdx.x, blockDim.x, fm_sub, Km_sub, Dm_sub);
    __syncthreads();   // this CTA wrote the blocks it reads next
    scatter_env(B, env, deriv, fm_sub, Km_sub, Dm_sub, fm, Km, Dm, threadIdx.x, blockDim.x);
}

void launch_finalize_pairs(const BatchDev& B, bool deriv, bool dm, double* fm_sub, double* Km_sub, double* Dm_sub,
                           cudaStream_t stream) {
    dim3 grid(B.n_pairs, B.env_count);
    k_reduce_items<<<grid, 128, 0, stream>>>(B, deriv ? 1 : 0, dm ? 1 : 0, fm_sub, Km_sub, Dm_sub);
}

void launch_reduce_scatter(const BatchDev& B, bool deriv, bool dm, double* fm_sub, double* Km_sub, double* Dm_sub, double* fm,
                           double* Km, double* Dm, cudaStream_t stream) {
    k_reduce_scatter<<<B.env_count, 256, 0, stream>>>(B, deriv ? 1 : 0, dm ? 1 : 0, fm_sub, Km_sub, Dm_sub, fm, Km, Dm);
}

void launch_scatter_system(const BatchDev& B, bool deriv, const double* fm_sub, const double* Km_sub,
                           const double* Dm_sub, double* fm, double* Km, double* Dm, cudaStream_t stream) {
    k_scatter_system<<<B.env_count, 256, 0, stream>>>(B, deriv ? 1 : 0, fm_sub, Km_sub, Dm_sub, fm, Km, Dm);
}

// --------------------------------------------------------------------------
// per-point query (original order) and raw sampling
// --------------------------------------------------------------------------
template <bool F32>
__device__ __forceinline__ void query_points_body(const PairDev& pd, const double* R1, const double* p1,
                                                  const double* R2, const double* p2, double* d_out, int32_t* cell_out,
                                                  uint8_t* active_out, int* s_count) {
    int cnt = 0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < pd.n; idx += gridDim.x * blockDim.x) {
        const double x0 = __ldg(pd.ox + idx), x1 = __ldg(pd.oy + idx), x2 = __ldg(pd.oz + idx);
        double xw[3], x[3];
        to_world(R1, p1, x0, x1, x2, xw);
        to_local(R2, p2, xw, x);
        const double d = __dadd_rn(query_sdf<F32>(pd.grid, x[0], x[1], x[2], false, nullptr), pd.grid.col_th);
        if (d_out) d_out[idx] = d;
        if (cell_out) {  // BodySDFObj::query_grid_location, BodySDFObj.cpp:379-383
            cell_out[3 * idx + 0] = (int)floor((x[0] - pd.grid.mn[0]) / pd.grid.dx);
            cell_out[3 * idx + 1] = (int)floor((x[1] - pd.grid.mn[1]) / pd.grid.dy);
            cell_out[3 * idx + 2] = (int)floor((x[2] - pd.grid.mn[2]) / pd.grid.dz);
        }
        const bool act = d < 0.;
        if (active_out) active_out[idx] = act ? 1 : 0;
        cnt += act ? 1 : 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(s_count, cnt);
}

__global__ void __launch_bounds__(256) k_query_points(PairDev pd, const double* __restrict__ E1,
                                                      const double* __restrict__ E2, double* d_out, int32_t* cell_out,
                                                      uint8_t* active_out, int32_t* n_active) {
    __shared__ double R1[9], p1[3], R2[9], p2[3];
    __shared__ int s_count;
    if (threadIdx.x < 12) stage_pose(E1, 1, 0, 0, R1, p1, threadIdx.x);
    else if (threadIdx.x >= 32 && threadIdx.x < 44) stage_pose(E2, 1, 0, 0, R2, p2, threadIdx.x - 32);
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    if (pd.grid.f32()) query_points_body<true>(pd, R1, p1, R2, p2, d_out, cell_out, active_out, &s_count);
    else query_points_body<false>(pd, R1, p1, R2, p2, d_out, cell_out, active_out, &s_count);
    __syncthreads();
    if (threadIdx.x == 0 && s_count) atomicAdd(n_active, s_count);
}

void launch_query_points(const PairDev& pd, const double* E1, const double* E2, double* d_out, int32_t* cell_out,
                         uint8_t* active_out, int32_t* n_active, cudaStream_t stream) {
    cudaMemsetAsync(n_active, 0, sizeof(int32_t), stream);
    const int blocks = max(1, min((pd.n + 255) / 256, 148 * 8));
    k_query_points<<<blocks, 256, 0, stream>>>(pd, E1, E2, d_out, cell_out, active_out, n_active);
}

template <bool F32>
__global__ void __launch_bounds__(256) k_sdf_distance(GridDev G, int64_t n, const double* __restrict__ x,
                                                      const double* __restrict__ y, const double* __restrict__ z,
                                                      double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = query_sdf<F32>(G, __ldg(x + i), __ldg(y + i), __ldg(z + i), false, nullptr);  // Mesh.cpp:143-189
}

void launch_sdf_distance(const GridDev& G, int64_t n, const double* x, const double* y, const double* z, double* out,
                         cudaStream_t stream) {
    if (n <= 0) return;
    const int blocks = (int)max((int64_t)1, min((n + 255) / 256, (int64_t)148 * 16));
    if (G.f32()) k_sdf_distance<true><<<blocks, 256, 0, stream>>>(G, n, x, y, z, out);
    else k_sdf_distance<false><<<blocks, 256, 0, stream>>>(G, n, x, y, z, out);
}

// query_SDF(x, true), query_dSDF and query_ddSDF at arbitrary points of the grid's frame
// (BodySDFObj.cpp:123-205): the raw FD gradient / Hessian the contact kernels consume, exposed
// so that rows a6 / a7 of the scope table can be compared with the reference bit for bit.
// fast != 0 runs the shared-stencil path of the planner kernel (sdf_stencil), else the
// sample-by-sample path of the general kernel (sdf_stencil_generic).
template <bool F32>
__global__ void __launch_bounds__(128) k_sdf_stencil(GridDev G, int64_t n, const double* __restrict__ x,
                                                     const double* __restrict__ y, const double* __restrict__ z, int fast,
                                                     double* __restrict__ acc, double* __restrict__ grad,
                                                     double* __restrict__ hess) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n + 31) / 32 * 32;   // sdf_stencil votes across the warp: whole warps iterate together
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        const int64_t j = i < n ? i : n - 1;
        const double p[3] = {__ldg(x + j), __ldg(y + j), __ldg(z + j)};
        double d, g[3], H[9];
        if (fast) sdf_stencil<F32>(G, p, d, g, H);
        else sdf_stencil_generic<F32>(G, p, d, g, H);
        if (i >= n) continue;
        if (acc) acc[i] = query_sdf<F32>(G, p[0], p[1], p[2], true, nullptr);
        if (grad) { grad[3 * i] = g[0]; grad[3 * i + 1] = g[1]; grad[3 * i + 2] = g[2]; }
        if (hess)
#pragma unroll
            for (int q = 0; q < 9; ++q) hess[9 * i + q] = H[q];
    }
}

void launch_sdf_stencil(const GridDev& G, int64_t n, const double* x, const double* y, const double* z, int fast,
                        double* acc, double* grad, double* hess, cudaStream_t stream) {
    if (n <= 0) return;
    const int blocks = (int)max((int64_t)1, min((n + 127) / 128, (int64_t)148 * 8));
    if (G.f32()) k_sdf_stencil<true><<<blocks, 128, 0, stream>>>(G, n, x, y, z, fast, acc, grad, hess);
    else k_sdf_stencil<false><<<blocks, 128, 0, stream>>>(G, n, x, y, z, fast, acc, grad, hess);
}

__global__ void k_aos_to_soa(const double* __restrict__ aos, int64_t n, double* __restrict__ x, double* __restrict__ y,
                             double* __restrict__ z) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        x[i] = aos[3 * i]; y[i] = aos[3 * i + 1]; z[i] = aos[3 * i + 2];
    }
}
void launch_aos_to_soa(const double* aos, int64_t n, double* x, double* y, double* z, cudaStream_t stream) {
    if (n <= 0) return;
    const int blocks = (int)max((int64_t)1, min((n + 255) / 256, (int64_t)148 * 16));
    k_aos_to_soa<<<blocks, 256, 0, stream>>>(aos, n, x, y, z);
}

// Builds the fp32 cell records (see GridDev::c32) from the plain fp32 grid.
__global__ void k_build_cells(const float* __restrict__ g, int ni, int nj, int nk, float4* __restrict__ out) {
    const size_t n = (size_t)ni * nj * nk;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % ni), j = (int)((idx / ni) % nj), k = (int)(idx / ((size_t)ni * nj));
        const int ip = (i == ni - 1) ? i : i + 1, jp = (j == nj - 1) ? j : j + 1, kp = (k == nk - 1) ? k : k + 1;
        auto at = [&](int a, int b, int c) { return g[(size_t)a + (size_t)ni * ((size_t)b + (size_t)nj * (size_t)c)]; };
        out[2 * idx] = make_float4(at(i, j, k), at(ip, j, k), at(i, jp, k), at(ip, jp, k));
        out[2 * idx + 1] = make_float4(at(i, j, kp), at(ip, j, kp), at(i, jp, kp), at(ip, jp, kp));
    }
}
void launch_build_cells(const float* g32, int ni, int nj, int nk, float4* out, cudaStream_t stream) {
    const size_t n = (size_t)ni * nj * nk;
    const int blocks = (int)max((size_t)1, min((n + 255) / 256, (size_t)148 * 16));
    k_build_cells<<<blocks, 256, 0, stream>>>(g32, ni, nj, nk, out);
}

// Builds the bricked fp32 layout (GridDev::b32) from the plain fp32 grid: one thread per brick node,
// node indices past n - 1 replicate the last layer.
__global__ void k_build_bricks(const float* __restrict__ g, int ni, int nj, int nk, int nbx, int nby, int nbz,
                               float* __restrict__ out) {
    const size_t n = (size_t)nbx * nby * nbz * 64;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int in = (int)(idx & 63);
        const size_t b = idx >> 6;
        const int bx = (int)(b % nbx), by = (int)((b / nbx) % nby), bz = (int)(b / ((size_t)nbx * nby));
        const int i = min(4 * bx + (in & 3), ni - 1), j = min(4 * by + ((in >> 2) & 3), nj - 1), k = min(4 * bz + (in >> 4), nk - 1);
        out[idx] = g[(size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * (size_t)k)];
    }
}
void launch_build_bricks(const float* g32, int ni, int nj, int nk, int nbx, int nby, int nbz, float* out, cudaStream_t stream) {
    const size_t n = (size_t)nbx * nby * nbz * 64;
    const int blocks = (int)max((size_t)1, min((n + 255) / 256, (size_t)148 * 16));
    k_build_bricks<<<blocks, 256, 0, stream>>>(g32, ni, nj, nk, nbx, nby, nbz, out);
}

// Squared Lipschitz bound of the trilinear interpolant (GridDev::lip): inside a cell each gradient
// component is a convex combination of the cell's four edge differences along that axis, so
// |grad|^2 <= sum_axis (largest |edge difference| / spacing)^2; the grid-wide bound is the
// maximum over cells. Upper neighbours are clamped like the sampler's (BodySDFObj.cpp:137-140).
// Non-negative doubles order like their bit patterns, hence the integer atomicMax (out[0]).
// out[1]: smallest node value in the two outermost node layers (GridDev::bmin) as an order-preserving
// key (ordered_key); the faces of [min_box, max_box] lie inside those layers (BodySDFObj.cpp:66-70).
__device__ __forceinline__ unsigned long long ordered_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void k_grid_lipschitz(const double* __restrict__ g, int ni, int nj, int nk, double rdx, double rdy, double rdz,
                                 unsigned long long* __restrict__ out) {
    const size_t n = (size_t)ni * nj * nk, stride = (size_t)gridDim.x * blockDim.x;
    double best = 0.0, bmin = 1e300;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        const int i = (int)(c % ni), j = (int)((c / ni) % nj), k = (int)(c / ((size_t)ni * nj));
        const int i1 = min(i + 1, ni - 1), j1 = min(j + 1, nj - 1), k1 = min(k + 1, nk - 1);
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
            v[q] = g[(size_t)((q & 1) ? i1 : i) + (size_t)ni * ((size_t)((q & 2) ? j1 : j) + (size_t)nj * ((q & 4) ? k1 : k))];
        const double ex = fmax(fmax(fabs(v[1] - v[0]), fabs(v[3] - v[2])), fmax(fabs(v[5] - v[4]), fabs(v[7] - v[6]))) * rdx;
        const double ey = fmax(fmax(fabs(v[2] - v[0]), fabs(v[3] - v[1])), fmax(fabs(v[6] - v[4]), fabs(v[7] - v[5]))) * rdy;
        const double ez = fmax(fmax(fabs(v[4] - v[0]), fabs(v[5] - v[1])), fmax(fabs(v[6] - v[2]), fabs(v[7] - v[3]))) * rdz;
        double l2 = ex * ex + ey * ey + ez * ez;
        bool finite = true;
#pragma unroll
        for (int q = 0; q < 8; ++q) finite = finite && isfinite(v[q]);
        if (!finite || !isfinite(l2)) l2 = 1e300;   // NaN / inf grids: never cull
        best = fmax(best, l2);
        if (i <= 1 || i >= ni - 2 || j <= 1 || j >= nj - 2 || k <= 1 || k >= nk - 2) bmin = finite ? fmin(bmin, v[0]) : -1e300;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
        bmin = fmin(bmin, __shfl_xor_sync(0xffffffffu, bmin, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (best > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(best));
        atomicMin(out + 1, ordered_key(bmin));
    }
}

void launch_grid_lipschitz(const double* g64, int ni, int nj, int nk, double dx, double dy, double dz,
                           unsigned long long* out, cudaStream_t stream) {
    const size_t n = (size_t)ni * nj * nk;
    const int blocks = (int)max((size_t)1, min((n + 255) / 256, (size_t)148 * 16));
    k_grid_lipschitz<<<blocks, 256, 0, stream>>>(g64, ni, nj, nk, 1.0 / dx, 1.0 / dy, 1.0 / dz, out);
}

// Self-test of exact_div(): n pseudo-random numerators (full 52-bit mantissas,
// magnitudes 2^-40 .. 2^12) divided by b, compared bit for bit with the IEEE `/`.
__global__ void k_selftest_exact_div(int64_t n, uint64_t seed, double b, double rb, unsigned long long* mismatches) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);  // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        const uint64_t mant = z & 0x000FFFFFFFFFFFFFull;
        const int e = (int)((z >> 52) % 53) - 40;  // exponent in [-40, 12]
        const double a = __longlong_as_double((long long)(((uint64_t)(1023 + e) << 52) | mant));
        const double q_ref = a / b;
        const double q = exact_div(a, b, rb);
        if (__double_as_longlong(q) != __double_as_longlong(q_ref)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}
void launch_selftest_exact_div(int64_t n, uint64_t seed, double b, unsigned long long* mismatches, cudaStream_t stream) {
    k_selftest_exact_div<<<148 * 8, 256, 0, stream>>>(n, seed, b, 1.0 / b, mismatches);
}

// min over n values: per-block minima into scratch[gridDim], then one block.
__global__ void k_min_stage(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
    __shared__ double sm[32];
    double m = DBL_MAX;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmin(m, v[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : DBL_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) out[blockIdx.x] = m;
    }
}
void launch_min_reduce(const double* v, int64_t n, double* out_scalar, double* scratch, cudaStream_t stream) {
    const int blocks = (int)max((int64_t)1, min((n + 255) / 256, (int64_t)1024));
    k_min_stage<<<blocks, 256, 0, stream>>>(v, n, scratch);
    k_min_stage<<<1, 256, 0, stream>>>(scratch, blocks, out_scalar);
}

}  // namespace sdfc
