// sdfc_host.h — host-side object definitions shared by the C-ABI translation units
// (sdfc_capi.cu: contact objects; sdfc_step.cu: batched BDF1 stepper). Internal.
#pragma once
#include "../../include/sdfcontact_b200.h"
#include "sdfc_kernels.cuh"

#include <cstdlib>
#include <string>
#include <vector>

namespace sdfc_host {
sdfc_status fail(sdfc_status code, const std::string& msg);
}
#define SDFC_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return sdfc_host::fail(SDFC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
#define SDFC_REQUIRE(cond, msg)                                        \
    do {                                                               \
        if (!(cond)) return sdfc_host::fail(SDFC_ERR_INVALID, msg);    \
    } while (0)

namespace sdfc_host {
template <class T>
sdfc_status dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) return SDFC_OK;
    SDFC_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    // SDFC_POISON=1 (debugging aid): every fresh allocation is filled with 0xFF bytes (NaN doubles / floats,
    // -1 ints), so that a read of memory the library never wrote shows up as NaN or a wild index instead
    // of depending on what the previous owner of the pages left there
    static const bool poison = [] { const char* e = std::getenv("SDFC_POISON"); return e && std::atoi(e) != 0; }();
    if (poison) { SDFC_CUDA(cudaMemset((void*)*p, 0xFF, count * sizeof(T))); SDFC_CUDA(cudaDeviceSynchronize()); }
    return SDFC_OK;
}
}  // namespace sdfc_host

struct sdfc_context {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    // optional per-kernel timing: (slot, start, stop) triples pending resolution
    bool timing = false;
    struct Span { int slot; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> free_events;
    double acc_ms[4] = {0, 0, 0, 0};
    int64_t acc_n[4] = {0, 0, 0, 0};

    cudaEvent_t get_event() {
        if (!free_events.empty()) { cudaEvent_t e = free_events.back(); free_events.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void begin(int slot) {
        if (!timing) return;
        Span s{slot, get_event(), get_event()};
        cudaEventRecord(s.a, stream);
        spans.push_back(s);
    }
    void end() {
        if (!timing) return;
        cudaEventRecord(spans.back().b, stream);
    }
};

struct sdfc_sdf {
    sdfc_context* ctx = nullptr;
    sdfc::GridDev dev{};
    double* d_g64 = nullptr;
    float4* d_c32 = nullptr;  // fp32 cell records (2 float4 per cell) when the grid is fp32-lossless
    float* d_b32 = nullptr;   // or the bricked fp32 grid (large grids), with its TMA descriptor table
    void* d_tmaps = nullptr;
    size_t count = 0;
    uint64_t version = 1;
    // grow-only scratch for the host-buffer sampling entry
    double* d_aos = nullptr; double* d_soa = nullptr; double* d_out = nullptr; double* d_red = nullptr;
    int64_t scratch_n = 0;
};

struct sdfc_points {
    sdfc_context* ctx = nullptr;
    int n = 0;
    double* d_o = nullptr;  // original order SoA: x[n] y[n] z[n]
    double* d_s = nullptr;  // sorted SoA (== d_o when not sorted)
    double* d_wsphere = nullptr;  // (cx, cy, cz, r) per 32-point word of the sorted copy
    double bsphere[4] = {0, 0, 0, 0};  // (cx, cy, cz, r) of the whole set
    double* d_tile_box = nullptr;  // 7 x (lo[3], hi[3]) per detection tile of the sorted copy: tile, halves, quarters
    std::vector<int32_t> perm;     // sorted position -> original index (empty: identity)
    bool sorted = false;
};

struct sdfc_pair {
    sdfc_context* ctx = nullptr;
    sdfc_points* pts = nullptr;
    sdfc_sdf* sdf = nullptr;
    double kn = 1, kt = 0, mu = 0, damping = 0, scale = 1, contact_scale = 1;
    uint64_t version = 1;
    sdfc_batch* single = nullptr;  // lazily created 1-env batch behind the per-pair entries
    // per-point query scratch
    double* d_d = nullptr; int32_t* d_cell = nullptr; uint8_t* d_active = nullptr; int32_t* d_count = nullptr;
    double* d_E = nullptr;  // 32 doubles
};

struct sdfc_ground_set;  // sdfc_ground.cu: E environments x G ground contacts

struct sdfc_batch {
    sdfc_context* ctx = nullptr;
    sdfc_ground_set* gset = nullptr;  // optional ForceGroundContact forces of every environment
    int n_envs = 0, n_bodies = 0, n_pairs = 0;
    std::vector<sdfc_pair*> pairs;
    std::vector<int32_t> body_a, body_b;
    std::vector<uint64_t> seen_pair_version, seen_sdf_version;
    sdfc::BatchDev dev{};
    sdfc::PairDev* d_pairs = nullptr;
    int* d_tile_pair = nullptr; int* d_tile_word0 = nullptr;
    uint32_t* d_mask = nullptr; int32_t* d_n_active = nullptr; int32_t* d_tile_count = nullptr;
    // work lists of the force kernels (sdfc::ForcePlanDev): slot 0 serves every evaluation on the context's stream;
    // the stepper takes one more per extra lane (lanes run concurrently on their own streams)
    static constexpr int kMaxPlanSlots = 4;
    sdfc::ForcePlanDev plan_slot[kMaxPlanSlots][2] = {};
    int n_plan_slots = 0;
    int plan_cap = 0;
    double* d_fm_sub = nullptr; double* d_Km_sub = nullptr; double* d_Dm_sub = nullptr;
    bool any_fric = false, any_damp = false;
    unsigned long long* d_brick_stats = nullptr;
    int64_t queries = 0;
    // staging for the host-buffer entry
    double* d_E = nullptr; double* d_phi = nullptr; double* d_fm = nullptr; double* d_Km = nullptr; double* d_Dm = nullptr;
    double* h_in = nullptr;   // pinned: E then phi
    double* h_out = nullptr;  // pinned: fm, Km, Dm
    int32_t* h_nact = nullptr;
    // the evaluation sequence of batch_run as a CUDA graph (captured at the second call with the same signature)
    static constexpr int kRunGraphs = 4;   // e.g. value-only and derivative evaluations alternate on the plug-in path
    struct RunGraph { cudaGraphExec_t exec = nullptr; uint64_t key = 0; int64_t launches = 0; uint64_t last_use = 0; };
    RunGraph run_graphs[kRunGraphs];
    uint64_t run_seen[kRunGraphs] = {0, 0, 0, 0};   // signatures seen once (launched plainly), next sighting captures
    uint64_t run_clock = 0;
    bool run_graph_ok = true;
    double* h_blocks = nullptr;   // pinned: per-pair / per-ground blocks of sdfc_batch_eval_blocks
    size_t h_blocks_bytes = 0;
    // optional temporal-coherence culling of the detection kernel (BatchDev::skin_*), switched on by the stepper
    bool skin_on = false;
    float* d_skin_margin = nullptr; double* d_skin_pose = nullptr; int* d_skin_int = nullptr;
    unsigned long long* d_skin_stats = nullptr;
};


namespace sdfc_host {
// Runs detection + force + per-pair finalisation (+ optional system scatter) of one batch.
// env_mode_dev == nullptr: every environment is evaluated, with derivatives iff deriv != 0.
// env_mode_dev != nullptr: per-environment 0 = skip, 1 = value only, 2 = with derivatives
// (both force variants are launched; CTAs of non-matching environments exit at once).
// env0 / env_count: environment range of this call (env_count < 0 = all environments).
sdfc_status batch_run(sdfc_batch* b, const double* E_dev, const double* phi_dev, int deriv, const int* env_mode_dev,
                      bool scatter, double* fm_dev, double* Km_dev, double* Dm_dev, int32_t* n_active_dev,
                      int env0 = 0, int env_count = -1, bool counts_zeroed = false);
// The contact kernels of one stepper round (sdfc_step.cu), fork / join over three streams so that the value-only
// and the derivative force kernels and the ground kernels run side by side: detection (s0) -> force<value> (s0)
// | force<deriv> (s1), ground contacts (s2) from the start. The round kernel has already zeroed the contact
// counters and run the temporal-coherence bookkeeping, and sums the split partials itself afterwards.
// ev: 4 events. Capturable into a CUDA graph.
sdfc_status batch_run_stepper(sdfc_batch* b, const double* E_dev, const double* phi_dev, const int* env_mode_dev, int env0,
                              int env_count, int plan_slot, cudaStream_t s0, cudaStream_t s1, cudaStream_t s2, cudaEvent_t* ev);
// Work-list buffers for `n` concurrent evaluations of the batch (stepper lanes); slot 0 exists from creation.
sdfc_status batch_ensure_plan_slots(sdfc_batch* b, int n);
// Uploads the pair table if parameters changed (batch_run does it on every call).
sdfc_status batch_prepare(sdfc_batch* b);
// Allocates the skin state of a batch and switches the culling on (idempotent).
sdfc_status batch_enable_skin(sdfc_batch* b);
}  // namespace sdfc_host

namespace sdfc_host {
// ForceGroundContact forces of a batch (sdfc_ground.cu). Outputs per (env, ground):
// 78 doubles = fm[6], Km[36], Dm[36] (6x6 body-diagonal blocks, column-major).
sdfc_status ground_set_create(sdfc_context* ctx, int n_envs, int n_bodies, int n, sdfc_ground* const* grounds,
                              const int32_t* body, sdfc_ground_set** out);
void ground_set_destroy(sdfc_ground_set* s);
// Uploads the ground table if parameters changed (ground_set_run does it on every call; separately callable so that
// the upload - which synchronises - happens before a stream capture).
sdfc_status ground_set_prepare(sdfc_ground_set* s);
sdfc_status ground_set_run(sdfc_ground_set* s, const double* E_dev, const double* phi_dev, int deriv,
                           const int* env_mode_dev, int env0, int env_count, cudaStream_t stream_override = nullptr,
                           bool counts_zeroed = false);
sdfc_status ground_set_scatter(sdfc_ground_set* s, int deriv, double* fm_dev, double* Km_dev, double* Dm_dev);
const double* ground_set_out(const sdfc_ground_set* s);
int32_t* ground_set_n_active(const sdfc_ground_set* s);
int ground_set_count(const sdfc_ground_set* s);
const int32_t* ground_set_bodies(const sdfc_ground_set* s);
}  // namespace sdfc_host
