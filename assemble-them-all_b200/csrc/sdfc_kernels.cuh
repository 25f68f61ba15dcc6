// sdfc_kernels.cuh — kernel-side data layout and launch entry points.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "sdfc_device.cuh"

namespace sdfc {

// Device descriptor of one contact pair (constant across environments).
struct PairDev {
    // contact points of the contact body, SoA. "s*" = spatially sorted copy used
    // by the force path and batched detection; "o*" = original order, used
    // whenever a per-point output is returned to the caller.
    const double *sx, *sy, *sz;
    const double *ox, *oy, *oz;
    // bounding sphere (cx, cy, cz, r) of every 32-point word of the sorted copy,
    // in the contact body's frame: lets detection skip whole words that cannot
    // reach the SDF's box (they would all return FLT_MAX, BodySDFObj.cpp:170)
    const double* wsphere;
    // axis-aligned boxes (lo[3], hi[3]) of every detection tile (kDetectTileWords words) of the sorted copy, in the
    // contact body's frame, 7 per tile: the tile, its two halves, its four quarters. The voxel bricks a tile (or
    // a part of it, when the whole does not fit) stages in shared memory are the grid region under the image of
    // the box (sdfc_kernels.cu, k_detect_brick)
    const double* tile_box;
    double bsphere[4];   // bounding sphere of the whole point set (pair-level culling of the stepper)
    int n;            // number of contact points
    int n_words;      // ceil(n / 32) mask words
    int word_offset;  // first mask word of this pair inside an environment's mask row
    int tile_offset;  // first detection tile of this pair inside the batch's tile list
    int body_a, body_b;  // pose slots inside an environment
    GridDev grid;
    PairParams prm;   // effective kn, kt (scaled), mu, damping
};

// Output value layout of one force item (ForcePlanDev) partial, fp64:
//   derivative: 4 K-blocks x kSlotK slots, then (if DM) 4 D-blocks x kSlotD slots
//   value-only: kSlotF slots
// Slot -> entry maps are produced by tr_index<N>() (see sdfc_kernels.cu).
#ifndef SDFC_FORCE_WARPS
#define SDFC_FORCE_WARPS 4
#endif
constexpr int kNumWarps = SDFC_FORCE_WARPS;   // force kernel: warps per CTA
constexpr int kForceThreads = 32 * kNumWarps;
constexpr int kDetectThreads = 256;     // widest detection CTA (plain and brick kernels)
constexpr int kDetectTileWords = 64;    // mask words (2048 points) per detection tile, whatever the CTA width
// The stepper's detection (k_detect_mask<true>) runs on narrower CTAs: after the margin test most tiles have a
// few live words, the kernel is a per-CTA latency chain (pair record -> poses / skin state -> broad phase -> live
// words) times the number of waves, and at 4 warps per CTA twice as many chains are in flight per SM
// (ncu: 42.8 % of the stalls were warps waiting at the CTA barriers while 2 of 8 warps ran the broad phase):
// 47.8 -> 45.2 ms per forward(100) of C2. The plain kernel gains 4 % the same way (0.1116 -> 0.1064 ms on C2;
// 64 threads: 0.117 ms); the brick kernel stays at 256 threads (3 CTAs per SM by shared memory).
#ifndef SDFC_DETECT_THREADS_SKIN
#define SDFC_DETECT_THREADS_SKIN 128
#endif
#ifndef SDFC_DETECT_THREADS_PLAIN
#define SDFC_DETECT_THREADS_PLAIN 128
#endif
constexpr int kDetectThreadsSkin = SDFC_DETECT_THREADS_SKIN;
constexpr int kDetectThreadsPlain = SDFC_DETECT_THREADS_PLAIN;

constexpr int tr_out(int n) {  // values left per lane after the 5 halving steps
    for (int s = 0; s < 5; ++s) n = (n + 1) / 2;
    return n;
}
constexpr int kOutK = tr_out(39);  // 2
constexpr int kOutD = tr_out(36);  // 2
constexpr int kOutF = tr_out(12);  // 1
constexpr int kPartialDeriv = 4 * kOutK * 32;         // 256 doubles
constexpr int kPartialDerivDM = kPartialDeriv + 4 * kOutD * 32;  // 512 doubles
constexpr int kPartialValue = kOutF * 32;             // 32 doubles
// planner fast path (derivatives, no friction, no damping): 12 Km columns + fm, 12 entries each, as two
// half-sums (contacts on lanes 0..15 / 16..31 of a round): column j, half h, entry r at (j * 2 + h) * 12 + r
constexpr int kNfCols = 13;
constexpr int kNfLanes = 24;                           // lanes that own one (half, entry) of every column
constexpr int kPartialNF = kNfCols * kNfLanes;         // 312 doubles

// Work list of one force launch. The active contacts of every (environment, pair) of the launch are cut into
// ITEMS of G x 32 consecutive contacts (set bits of the pair's mask row, in word order); the items of a pair are
// consecutive. Written by k_force_plan after detection, consumed by k_force (one warp per item, taken from the
// `next` counter) and by reduce_items_one (fixed-order sum of a pair's item partials).
struct ForcePlanDev {
    // [cap][2]  entry 0: x = env * n_pairs + pair, y / z = the item's contacts [y, z) in mask order (z - y <= 32 G),
    //                     w = contacts of the pair in front of the detection tile that holds contact y;
    //           entry 1: x = index in `mask` of that tile's first word, y = mask words of the pair from there on,
    //                     z / w = pose slots env * n_bodies + body_a / body_b
    int4* items;
    int* pair_item0;   // [n_envs * n_pairs]  first item of the pair (meaningful while it has active contacts)
    // [4]  0: number of items (handed out by atomics while planning), 1: next item to evaluate, 2: upper bound of the
    // 32-contact rounds of the evaluation (detection adds every tile's share to plan[0]'s), 3: G. 0..2 are cleared
    // before every detection (launch_detect_mask / the stepper's round kernel), 3 always holds a valid G.
    int* hdr;
    double* partial;   // [cap][stride]  one slot per item
    int cap;           // item capacity: G doubles until the list fits
    int stride;        // doubles per slot
    int g_min;         // smallest G (1; SDFC_FORCE_G)
};

struct BatchDev {
    const PairDev* pairs;  // [n_pairs]
    int n_pairs, n_bodies, n_envs;
    int words_per_env;
    // detection tiles: tile t covers kDetectTileWords words of pair tile_pair[t]
    const int* tile_pair;
    const int* tile_word0;
    int n_tiles;
    int any_bricks;     // some pair samples a bricked grid (GridDev::b32): detection stages voxel bricks with TMA
    int brick_cap;      // shared-memory bytes a tile's brick may take (larger bricks: global gathers); SDFC_BRICK_SMEM
    unsigned long long* brick_stats;   // [3] tiles staged / too large / outside the grid box (diagnostics), or nullptr
    uint32_t* mask;     // [n_envs][words_per_env]
    int32_t* n_active;  // [n_envs][n_pairs]
    int32_t* tile_count;   // [n_envs][n_tiles]  active contacts per detection tile (written by every tile's CTA)
    ForcePlanDev plan[2];   // work lists of the value-only (0) and the derivative (1) force kernel
    // optional per-environment evaluation mode (batched stepper): nullptr = every
    // environment, derivative flag taken from the launch; else 0 = skip,
    // 1 = value only, 2 = with derivatives
    const int* env_mode;
    // environment range of this launch: [env0, env0 + env_count). The stepper runs
    // disjoint ranges ("lanes") on separate streams so that one lane's latency-bound
    // assembly / solve kernels overlap another lane's contact kernels.
    int env0, env_count;
    // Optional temporal-coherence state of the stepper ("skin", nullptr = off). For every
    // (env, pair) a reference relative pose of the contact body in the SDF frame, and for every
    // 32-point word the smallest d = sdf + col_th its points had at that pose (or -1 when a point
    // was outside the grid box). While L * (largest displacement of the word since the
    // reference) stays below that margin no point of the word can have d < 0, so the word is
    // skipped with an empty mask - exactly what a full evaluation would produce.
    float* skin_margin;   // [n_envs][words_per_env]
    double* skin_ref;     // [n_envs][n_pairs][12]  R_rel (row-major) | p_rel at the last rebuild
    double* skin_cur;     // [n_envs][n_pairs][12]  same for the current evaluation
    int* skin_mode;       // [n_envs][n_pairs]      1 = rebuild margins in this evaluation, 2 = whole pair skipped
    int* skin_pair_margin;   // [n_envs][n_pairs]   float bits: smallest word margin (-1 as soon as one word has none)
    int* skin_valid;      // [n_envs][n_pairs]      reference pose and margins exist
    int* skin_fail;       // [n_envs][n_pairs]      eligible words that had to be evaluated last time
    int* skin_elig;       // [n_envs][n_pairs]      words with a margin
    unsigned long long* skin_stats;   // [4] running totals: words culled, words whose margin was used up, rebuilds, pairs skipped
};

// Skin bookkeeping of one (environment, pair) before a detection of the stepper: the relative pose now,
// and whether this evaluation rebuilds the margins (mode 1), skips the whole pair (mode 2) or tests words
// against their margins (mode 0). Called by k_skin_decide and, fused, by the stepper's round kernel.
__device__ __forceinline__ void skin_decide_one(const BatchDev& B, const double* E, int env, int pair) {
    const size_t ep = (size_t)env * B.n_pairs + pair;
    const PairDev& pd = B.pairs[pair];
    const double* E1 = E + ((size_t)env * B.n_bodies + pd.body_a) * 16;
    const double* E2 = E + ((size_t)env * B.n_bodies + pd.body_b) * 16;
    double cur[12];
    for (int r = 0; r < 3; ++r) {          // R_rel = R2^T R1, p_rel = R2^T (p1 - p2); column-major poses
        for (int c = 0; c < 3; ++c)
            cur[3 * r + c] = E2[4 * r] * E1[4 * c] + E2[4 * r + 1] * E1[4 * c + 1] + E2[4 * r + 2] * E1[4 * c + 2];
        cur[9 + r] = E2[4 * r] * (E1[12] - E2[12]) + E2[4 * r + 1] * (E1[13] - E2[13]) + E2[4 * r + 2] * (E1[14] - E2[14]);
    }
    int mode = 0;
    if (!B.skin_valid[ep]) mode = 1;
    else if (4 * B.skin_fail[ep] > B.skin_elig[ep]) mode = 1;   // > 25 % of the margins are used up
    else {
        // whole pair: every word has a margin and the smallest one outlasts the largest displacement of the set
        const float pm = __int_as_float(B.skin_pair_margin[ep]);
        if (pm >= 0.f) {
            const double* ref = B.skin_ref + ep * 12;
            double d[12], f2 = 0.0;
            for (int i = 0; i < 12; ++i) { d[i] = cur[i] - ref[i]; if (i < 9) f2 += d[i] * d[i]; }
            const double cx = pd.bsphere[0], cy = pd.bsphere[1], cz = pd.bsphere[2];
            const double ux = d[0] * cx + d[1] * cy + d[2] * cz + d[9];
            const double uy = d[3] * cx + d[4] * cy + d[5] * cz + d[10];
            const double uz = d[6] * cx + d[7] * cy + d[8] * cz + d[11];
            const double delta = sqrt(ux * ux + uy * uy + uz * uz) + sqrt(f2) * pd.bsphere[3];
            if ((double)pm - pd.grid.lip * delta * (1.0 + 1e-9) > 1e-9) { mode = 2; atomicAdd(B.skin_stats + 3, 1ull); }
        }
    }
    B.skin_fail[ep] = 0;
    for (int i = 0; i < 12; ++i) B.skin_cur[ep * 12 + i] = cur[i];
    if (mode == 1) {
        for (int i = 0; i < 12; ++i) B.skin_ref[ep * 12 + i] = cur[i];
        B.skin_valid[ep] = 1;
        B.skin_elig[ep] = 0;
        B.skin_pair_margin[ep] = __float_as_int(3e38f);
        atomicAdd(B.skin_stats + 2, 1ull);
    }
    B.skin_mode[ep] = mode;
}

// launches (all asynchronous on `stream`)
// reset_counts = false: the caller has already zeroed n_active of the environments it evaluates (the stepper)
void launch_detect_mask(const BatchDev& B, const double* E, cudaStream_t stream, bool reset_counts = true);
// skin bookkeeping of one evaluation (before launch_detect_mask when B.skin_margin != nullptr)
void launch_skin_decide(const BatchDev& B, const double* E, cudaStream_t stream);
void launch_force(const BatchDev& B, const double* E, const double* phi, bool deriv, bool any_fric, bool any_damp,
                  cudaStream_t stream);
// Work lists of the force kernels from the contact counts of a finished detection (one warp per (environment,
// pair)). B.env_mode == nullptr: the list `deriv` for every environment of the launch; else both lists,
// environments sorted by their mode.
void launch_force_plan(const BatchDev& B, bool deriv, cudaStream_t stream);
int device_sm_count();
// Sums the item partials in fixed order into per-pair blocks (overwrite):
// fm_sub [env][pair][12], Km_sub / Dm_sub [env][pair][144] column-major.
void launch_finalize_pairs(const BatchDev& B, bool deriv, bool dm, double* fm_sub, double* Km_sub, double* Dm_sub,
                           cudaStream_t stream);
// Gathers the pair blocks into per-environment system arrays in force order
// (overwrite): fm [env][6*n_bodies], Km / Dm [env][(6*n_bodies)^2] column-major.
// Dm may be nullptr.
// launch_finalize_pairs + launch_scatter_system in one kernel (a CTA per environment); B.env_mode must be nullptr
void launch_reduce_scatter(const BatchDev& B, bool deriv, bool dm, double* fm_sub, double* Km_sub, double* Dm_sub, double* fm,
                           double* Km, double* Dm, cudaStream_t stream);
void launch_scatter_system(const BatchDev& B, bool deriv, const double* fm_sub, const double* Km_sub,
                           const double* Dm_sub, double* fm, double* Km, double* Dm, cudaStream_t stream);
// Per-point outputs in ORIGINAL order for one pair, one environment.
void launch_query_points(const PairDev& pair_host_copy, const double* E1, const double* E2, double* d_out,
                         int32_t* cell_out, uint8_t* active_out, int32_t* n_active, cudaStream_t stream);
// Raw SDF sampling (SDFMesh::distance): SoA points in the grid's frame.
void launch_sdf_distance(const GridDev& G, int64_t n, const double* x, const double* y, const double* z, double* out,
                         cudaStream_t stream);
// query_SDF(x, true) / query_dSDF / query_ddSDF at SoA points of the grid's frame (any output may be nullptr).
void launch_sdf_stencil(const GridDev& G, int64_t n, const double* x, const double* y, const double* z, int fast,
                        double* acc, double* grad, double* hess, cudaStream_t stream);
void launch_min_reduce(const double* v, int64_t n, double* out_scalar, double* scratch, cudaStream_t stream);
void launch_build_cells(const float* g32, int ni, int nj, int nk, float4* out, cudaStream_t stream);
// plain fp32 grid -> 4 x 4 x 4 bricks with one replicated node layer past n - 1 (GridDev::b32)
void launch_build_bricks(const float* g32, int ni, int nj, int nk, int nbx, int nby, int nbz, float* out, cudaStream_t stream);
// shared-memory bytes a k_detect_brick CTA may use for its voxel brick
constexpr int kBrickSmemBytes = 72 * 1024;   // default BatchDev::brick_cap (3 CTAs / SM)
constexpr int kBrickSmemMax = 200 * 1024;
constexpr int kBrickMaxXY = 16;   // tensor-map table: box = 64 floats x (1..16) x (1..16) x 1 bricks
// *out (zeroed by the caller) = bit pattern of the squared Lipschitz bound of the interpolant (GridDev::lip)
void launch_grid_lipschitz(const double* g64, int ni, int nj, int nk, double dx, double dy, double dz,
                           unsigned long long* out, cudaStream_t stream);
void launch_selftest_exact_div(int64_t n, uint64_t seed, double b, unsigned long long* mismatches, cudaStream_t stream);
void launch_aos_to_soa(const double* aos, int64_t n, double* x, double* y, double* z, cudaStream_t stream);

}  // namespace sdfc
