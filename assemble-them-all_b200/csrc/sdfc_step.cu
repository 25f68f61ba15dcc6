// sdfc_step.cu — device-resident batched BDF1 stepper: the
// caller side of the contact path for E environments at once, so that
// Simulation::forward(n) never returns to the host between substeps.
//
// One evaluation "round" for all environments that still have work:
//   k_sim_pre    candidate state -> body poses E_0i and twists phi (joint + body
//                kinematics), per-environment evaluation mode (value / derivative)
//   batch_run    detection + contact force (+ Jacobians) kernels of sdfc_kernels.cu
//   k_sim_post   body forces, reduced assembly, BDF1 residual g (and H), LU solve,
//                line-search / Newton / step state machine
// Every environment advances through its own Newton iterations and line-search
// trials; the host only polls one counter (environments not yet finished).
//
// Restated from (paths relative to simulation/redmax/ of the reference):
//   Joint/Joint.cpp:95-141, JointFixed.cpp:6-13, JointTranslational.cpp:9-40,
//   JointFree3DExp.cpp:13-110, JointSphericalExp.cpp:13-183,237-264,
//   Body/Body.cpp:138-181,250-330, Robot.cpp:656-744 (root joints),
//   Simulation.cpp:379-503 (computeMatrices), :999-1079 (newton),
//   :1081-1112 (BDF1 residual), :914-966,1248-1257 (forward / integration_BDF1).
// One thread per environment: everything here is O(100) - O(10^4) flops per
// evaluation against O(10^7) in the contact kernels.
#include "sdfc_host.h"
#include "sdfc_force.cuh"
#include "sdfc_step_math.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

#ifndef SDFC_SKIN_DEFAULT
#define SDFC_SKIN_DEFAULT 1
#endif
#ifndef SDFC_TAIL_TEAM
#define SDFC_TAIL_TEAM 128
#endif

namespace sdfc {

constexpr int kCapR = kStepCapR;      // reduced dofs per environment
constexpr int kCapDyn = kStepCapDyn;  // non-fixed bodies per environment
constexpr double kNewtonTol = 1e-8;  // Simulation.cpp:1008
constexpr int kMaxLS = 20;           // Simulation.cpp:1012
constexpr int kMaxStrike = 3;        // Simulation.cpp:1014
constexpr int kSimWarps = 4;         // environments (warps) per CTA of k_sim_round

struct SimDev {
    const BodyDev* bodies;
    const int *pair_a, *pair_b;
    int n_bodies, n_pairs, n_envs, nr, n_dyn, max_iter;
    double h, grav[3];
    double *q, *qdot, *q0, *qdot0, *x, *dx;  // [n_envs][nr]
    double *gnorm, *alpha;                   // [n_envs]
    int *phase, *trial, *iter, *strike, *steps_left, *step_idx;
    uint8_t *converged, *touched;
    const uint8_t* active;                   // [n_envs] or nullptr (all environments step)
    const double* fext;                      // [n_envs][n_bodies][6]
    double *E, *phi;                         // contact-path inputs
    int* env_mode;
    double *fm_sub, *Km_sub, *Dm_sub;        // contact-path outputs per (env, pair): summed here from the force kernels' split partials
    int32_t* n_active;                       // [n_envs][n_pairs]; zeroed here for the next detection
    BatchDev B;                              // the contact batch (split partials, temporal-coherence state)
    int dm;                                  // some pair has friction or damping (layout of the partials)
    int n_grounds;                           // ForceGroundContact forces (0 = none)
    const int* ground_body;                  // [n_grounds]
    const double* g_out;                     // [n_envs][n_grounds][78]: fm 6, Km 36, Dm 36 (column-major)
    int32_t* g_n_active;                     // [n_envs][n_grounds]; zeroed here for the next evaluation
    uint8_t* touched_ground;                 // [n_envs][n_grounds]
    double *q_traj, *qdot_traj;              // [steps][n_envs][nr] or nullptr
    int* n_unfinished;
    unsigned long long* counters;            // deriv evals, value evals, steps
    int env0, env_count;                     // environment range of this launch (a stepper lane)
};

enum { PH_DERIV = 0, PH_LS = 1, PH_DONE = 2 };
enum { ROUND_FULL = 0, ROUND_FIRST = 1, ROUND_POSES = 2 };

// ------------------------------------------------------------------ kernels
__global__ void k_sim_begin(SimDev S, int n_steps) {
    const int env = S.env0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= S.env0 + S.env_count) return;
    // environments masked out by sdfc_sim_set_active keep their state untouched
    const bool on = n_steps > 0 && (S.active == nullptr || S.active[env] != 0);
    if (on) atomicAdd(S.n_unfinished, 1);
    S.steps_left[env] = on ? n_steps : 0;
    S.step_idx[env] = 0;
    S.phase[env] = on ? PH_DERIV : PH_DONE;
    S.iter[env] = 0;
    S.strike[env] = 0;
    const size_t o = (size_t)env * S.nr;
    for (int j = 0; j < S.nr; ++j) {  // integration_BDF1: q1 = q0 + h qdot0 (Simulation.cpp:1255)
        S.q0[o + j] = S.q[o + j];
        S.qdot0[o + j] = S.qdot[o + j];
        S.x[o + j] = S.q[o + j] + S.h * S.qdot[o + j];
    }
}

// candidate state of this round into the workspace: x (Newton evaluation) or x + alpha dx (line-search trial)
__device__ __forceinline__ void w_candidate(const WarpExec& wx, const SimDev& S, int env, int phase, EnvW& W) {
    const size_t o = (size_t)env * S.nr;
    const double alpha = S.alpha[env];
    WFOR(j, S.nr) {
        const double x = (phase == PH_LS) ? S.x[o + j] + alpha * S.dx[o + j] : S.x[o + j];
        W.xc[j] = x;
        W.qdc[j] = (x - S.q0[o + j]) / S.h;  // Simulation.cpp:1082,1095
    }
    WSYNC();
}

// One round of one environment per TEAM (TEAM = 32: a warp, kSimWarps environments per CTA; TEAM > 32: the whole
// CTA works on one environment - the tail of a forward, where a handful of environments are still iterating and
// the latency of one round of one environment is the whole cost; same statements, same operands, same bits).
// Workspace in shared memory, sdfc_step_math.cuh:
//   1. "post": the contact kernels have evaluated the candidate of the previous round; assemble the BDF1
//      residual g (and H), solve, advance the Newton / line-search / step state machine
//      (Simulation.cpp:999-1079, 955-966, 1259);
//   2. "pre": poses E_0i and twists phi of the next candidate for the contact kernels, its evaluation mode,
//      and the zeroed contact counters of the next detection.
// ROUND_FIRST skips 1 (first round of a forward), ROUND_POSES only writes the poses of the current state.
template <int TEAM>
__global__ void __launch_bounds__(TEAM == 32 ? kSimWarps * 32 : TEAM) k_sim_round(SimDev S, int kind) {
    extern __shared__ __align__(16) unsigned char ws_raw[];
    const int warp = (TEAM == 32) ? (threadIdx.x >> 5) : 0;
    const int env = S.env0 + ((TEAM == 32) ? blockIdx.x * kSimWarps + warp : blockIdx.x);
    if (env >= S.env0 + S.env_count) return;   // whole teams leave together
    EnvW& W = reinterpret_cast<EnvW*>(ws_raw)[warp];
    const WarpExec wx{(int)(TEAM == 32 ? (threadIdx.x & 31) : threadIdx.x), 1, TEAM};
    const int nr = S.nr;
    const double h = S.h;
    const size_t o = (size_t)env * nr;

    if (kind == ROUND_POSES) {
        WFOR(j, nr) { W.xc[j] = S.q[o + j]; W.qdc[j] = S.qdot[o + j]; }
        WSYNC();
        w_poses(wx, S.bodies, S.n_bodies, W, S.E + (size_t)env * S.n_bodies * 16, S.phi + (size_t)env * S.n_bodies * 6);
        return;
    }

    // work lists of this round's force kernels: empty (detection and k_force_plan fill them)
    if (blockIdx.x == 0 && threadIdx.x < 6) S.B.plan[threadIdx.x / 3].hdr[threadIdx.x % 3] = 0;
    const int mode = (kind == ROUND_FULL) ? S.env_mode[env] : 0;
    if (mode != 0) {
        const bool deriv = (mode == 2);
        w_candidate(wx, S, env, deriv ? PH_DERIV : PH_LS, W);
        if (WLANE0) atomicAdd(S.counters + (deriv ? 0 : 1), 1ull);
        // contact history (CollisionDetection.cpp:143-146, :67)
        WFOR(p, S.n_pairs)
            if (S.n_active[(size_t)env * S.n_pairs + p] > 0) S.touched[(size_t)env * S.n_pairs + p] = 1;
        WFOR(gi, S.n_grounds)
            if (S.g_n_active[(size_t)env * S.n_grounds + gi] > 0) S.touched_ground[(size_t)env * S.n_grounds + gi] = 1;
        // item partials of the force kernels -> per-pair blocks (what k_reduce_items does outside the stepper)
        for (int p = 0; p < S.n_pairs; ++p)
            reduce_items_one(S.B, env, p, deriv ? 1 : 0, S.dm, wx.lane, TEAM, S.fm_sub, S.Km_sub, S.Dm_sub);
        WSYNC();
        EvalIn in;
        in.bodies = S.bodies; in.n_bodies = S.n_bodies; in.n_pairs = S.n_pairs; in.n_dyn = S.n_dyn; in.nr = nr;
        in.n_grounds = S.n_grounds; in.pair_a = S.pair_a; in.pair_b = S.pair_b; in.ground_body = S.ground_body;
        in.h = h; in.grav[0] = S.grav[0]; in.grav[1] = S.grav[1]; in.grav[2] = S.grav[2];
        in.q0 = S.q0 + o; in.qdot0 = S.qdot0 + o;
        in.fext = S.fext + (size_t)env * S.n_bodies * 6;
        in.fm_sub = S.fm_sub + (size_t)env * S.n_pairs * 12;
        in.Km_sub = S.Km_sub + (size_t)env * S.n_pairs * 144;
        in.Dm_sub = S.Dm_sub + (size_t)env * S.n_pairs * 144;
        in.g_out = S.g_out + (size_t)env * S.n_grounds * 78;
        w_evaluate(wx, in, deriv, W);

        double gn = 0;
        for (int i = 0; i < nr; ++i) gn += W.g[i] * W.g[i];   // every lane: same operands, same order
        gn = sqrt(gn);

        // ---- Newton / line-search state machine (Simulation.cpp:999-1079)
        if (deriv) {
            WFOR(i, nr) W.rhs[i] = -W.g[i];
            WSYNC();
            w_lu_solve(wx, W.H, W.rhs, nr, &W.piv);
            WFOR(i, nr) S.dx[o + i] = W.rhs[i];
            if (WLANE0) {
                S.gnorm[env] = gn;
                S.alpha[env] = 1.0;
                S.trial[env] = 0;
                S.iter[env] += 1;
                S.phase[env] = PH_LS;
            }
        } else {
            // decisions from scalars every lane reads before lane 0 rewrites them
            const double gnorm = S.gnorm[env], alpha = S.alpha[env];
            const int trial0 = S.trial[env], strike0 = S.strike[env], iter0 = S.iter[env];
            WSYNC();
            bool accept = false, end_newton = false, success = false, next_trial = false;
            int strike = strike0;
            double alpha_use = alpha;
            if (gn < gnorm) {  // success_ls
                strike = 0;
                accept = true;
            } else {
                alpha_use = alpha * 0.5;
                if (trial0 + 1 < kMaxLS) next_trial = true;   // next trial: phase stays PH_LS
                else {
                    strike = strike0 + 1;
                    if (strike >= kMaxStrike) end_newton = true;  // break before applying the step
                    else accept = true;                            // applied with alpha = 0.5^20
                }
            }
            if (WLANE0) {
                S.strike[env] = strike;
                if (!(gn < gnorm)) { S.trial[env] = trial0 + 1; S.alpha[env] = alpha_use; }
            }
            if (!next_trial) {
                if (accept) {
                    WFOR(i, nr) S.x[o + i] += alpha_use * S.dx[o + i];
                    if (gn < kNewtonTol) { success = true; end_newton = true; }
                    else if (iter0 >= S.max_iter) end_newton = true;
                    else if (WLANE0) S.phase[env] = PH_DERIV;
                }
                WSYNC();
                if (end_newton) {
                    // ---- end of the step (Simulation.cpp:955-966, 1259): qdot1, reparam, history
                    WFOR(i, nr) { const double q1 = S.x[o + i]; W.xc[i] = q1; W.qdc[i] = (q1 - S.q0[o + i]) / h; }
                    WSYNC();
                    for (int b = 0; b < S.n_bodies; ++b) {
                        const BodyDev& B = S.bodies[b];
                        if (B.type != kJointFree3DExp) continue;
                        w_reparam(wx, W.xc + B.dof_off + 3, W.qdc + B.dof_off + 3, W);   // JointSphericalExp.cpp:237-264
                    }
                    const int step = S.step_idx[env];
                    const int left = S.steps_left[env] - 1;
                    WSYNC();
                    WFOR(i, nr) {
                        const double q1 = W.xc[i], qd1 = W.qdc[i];
                        S.q[o + i] = q1;
                        S.qdot[o + i] = qd1;
                        if (S.q_traj) S.q_traj[((size_t)step * S.n_envs + env) * nr + i] = q1;
                        if (S.qdot_traj) S.qdot_traj[((size_t)step * S.n_envs + env) * nr + i] = qd1;
                        if (left > 0) {
                            S.q0[o + i] = q1;
                            S.qdot0[o + i] = qd1;
                            S.x[o + i] = q1 + h * qd1;
                        }
                    }
                    if (WLANE0) {
                        S.converged[env] = success ? 1 : 0;
                        atomicAdd(S.counters + 2, 1ull);
                        S.step_idx[env] = step + 1;
                        S.steps_left[env] = left;
                        if (left <= 0) {
                            S.phase[env] = PH_DONE;
                            atomicSub(S.n_unfinished, 1);
                        } else {
                            S.iter[env] = 0;
                            S.strike[env] = 0;
                            S.phase[env] = PH_DERIV;
                        }
                    }
                }
            }
        }
        WSYNC();   // the state written above is read back below by other lanes
    }

    // ---- next candidate: evaluation mode, poses and twists for the contact kernels
    const int phase = S.phase[env];
    if (WLANE0) S.env_mode[env] = (phase == PH_DONE) ? 0 : (phase == PH_LS ? 1 : 2);
    if (phase == PH_DONE) return;
    WFOR(p, S.n_pairs) S.n_active[(size_t)env * S.n_pairs + p] = 0;   // the detection kernel accumulates into it
    WFOR(gi, S.n_grounds) S.g_n_active[(size_t)env * S.n_grounds + gi] = 0;
    w_candidate(wx, S, env, phase, W);
    w_poses(wx, S.bodies, S.n_bodies, W, S.E + (size_t)env * S.n_bodies * 16, S.phi + (size_t)env * S.n_bodies * 6);
    // temporal-coherence bookkeeping of the next detection (k_skin_decide outside the stepper)
    if (S.B.skin_margin != nullptr) WFOR(p, S.n_pairs) skin_decide_one(S.B, S.E, env, p);
}

}  // namespace sdfc

// ============================================================================
// host side of the sdfc_sim_* entry points
// ============================================================================
using namespace sdfc;
using sdfc_host::dev_alloc;
using sdfc_host::fail;

struct sdfc_sim {
    sdfc_batch* batch = nullptr;
    sdfc_context* ctx = nullptr;
    SimDev dev{};
    int n_envs = 0, n_bodies = 0, n_pairs = 0, nr = 0;
    BodyDev* d_bodies = nullptr;
    int *d_pair_a = nullptr, *d_pair_b = nullptr;
    double* d_state = nullptr;   // q, qdot, q0, qdot0, x, dx
    double* d_scalars = nullptr; // gnorm, alpha
    int* d_ints = nullptr;       // phase, trial, iter, strike, steps_left, step_idx, env_mode
    uint8_t *d_converged = nullptr, *d_touched = nullptr, *d_active = nullptr, *d_touched_ground = nullptr;
    int* d_ground_body = nullptr;
    int n_grounds = 0;
    // lanes: disjoint environment ranges stepped on their own streams (see BatchDev::env0)
    static constexpr int kMaxLanes = 4;
    int n_lanes = 1;
    cudaStream_t lane_stream[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t lane_event[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork_event = nullptr;
    // fork / join streams and events of a lane's contact kernels, and the captured steady-state round
    cudaStream_t lane_aux[kMaxLanes][2] = {};
    cudaEvent_t lane_ev[kMaxLanes][4] = {};
    cudaGraphExec_t lane_graph[kMaxLanes][3] = {};   // [lane][round-kernel variant: 0 warp, 1 / 2 CTA of 128 / 256 threads per environment]
    uint64_t lane_graph_key[kMaxLanes][3] = {};
    int tail_envs = 0;   // a lane switches to the CTA-per-environment variants once this few of its environments are unfinished
    int team_variant(int unfinished) const {   // 256 threads while one CTA per SM covers them, 128 up to tail_envs, else a warp
        if (unfinished > tail_envs) return 0;
        return unfinished <= tail_envs / 4 ? 2 : 1;
    }
    bool use_graph = true;
    bool use_active = false;
    double* d_fext = nullptr;
    int* d_unfinished = nullptr;
    unsigned long long* d_counters = nullptr;
    double *d_traj_q = nullptr, *d_traj_qd = nullptr;
    size_t traj_cap = 0;  // steps
    int* h_unfinished = nullptr;  // pinned
    int64_t rounds = 0;
};

namespace {
constexpr int kTailTeam = SDFC_TAIL_TEAM;   // threads per environment of the CTA-per-environment variants: this and twice this

sdfc_status launch_round(sdfc_sim* s, const SimDev& D, int kind, cudaStream_t st, int variant = 0) {
    const size_t smem = sizeof(EnvW) * kSimWarps;
    static bool configured = false;
    if (!configured) {
        SDFC_CUDA(cudaFuncSetAttribute(k_sim_round<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SDFC_CUDA(cudaFuncSetAttribute(k_sim_round<kTailTeam>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EnvW)));
        SDFC_CUDA(cudaFuncSetAttribute(k_sim_round<2 * kTailTeam>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EnvW)));
        configured = true;
    }
    if (variant == 2) {
        k_sim_round<2 * kTailTeam><<<D.env_count, 2 * kTailTeam, sizeof(EnvW), st>>>(D, kind);
    } else if (variant == 1) {
        k_sim_round<kTailTeam><<<D.env_count, kTailTeam, sizeof(EnvW), st>>>(D, kind);
    } else {
        const int blocks = (D.env_count + kSimWarps - 1) / kSimWarps;
        k_sim_round<32><<<blocks, kSimWarps * 32, smem, st>>>(D, kind);
    }
    s->ctx->launches += 1;
    return SDFC_OK;
}
}  // namespace

extern "C" {

sdfc_status sdfc_sim_create(sdfc_batch* batch, const sdfc_body_desc* bodies, const double gravity[3], double h,
                            sdfc_sim** out) {
    SDFC_REQUIRE(batch && bodies && gravity && out, "sdfc_sim_create: NULL argument");
    SDFC_REQUIRE(h > 0, "sdfc_sim_create: timestep must be positive");
    *out = nullptr;
    const int nb = batch->n_bodies, ne = batch->n_envs, np = batch->n_pairs;
    std::vector<BodyDev> hb(nb);
    int nr = 0, nd = 0;
    for (int b = 0; b < nb; ++b) {
        BodyDev& B = hb[b];
        const sdfc_body_desc& d = bodies[b];
        SDFC_REQUIRE(d.joint_type == SDFC_JOINT_FIXED || d.joint_type == SDFC_JOINT_TRANSLATIONAL ||
                     d.joint_type == SDFC_JOINT_FREE3D_EXP, "sdfc_sim_create: unsupported joint type");
        make_body_dev(d.joint_type, d.E_pj_0, d.E_ji, d.inertia, d.mass, nr, nd, B);
        nr += B.ndof;
        if (B.ndof > 0) ++nd;
    }
    SDFC_REQUIRE(nr <= kCapR && nd <= kCapDyn, "sdfc_sim_create: at most 2 non-fixed bodies / 12 reduced dofs are supported");
    SDFC_REQUIRE(nr > 0, "sdfc_sim_create: no degrees of freedom");
    SDFC_CUDA(cudaSetDevice(batch->ctx->device));
    sdfc_sim* s = new sdfc_sim();
    s->batch = batch; s->ctx = batch->ctx; s->n_envs = ne; s->n_bodies = nb; s->n_pairs = np; s->nr = nr;
    sdfc_status st = SDFC_OK;
    auto A = [&](sdfc_status x) { if (st == SDFC_OK) st = x; };
    A(dev_alloc(&s->d_bodies, (size_t)nb));
    A(dev_alloc(&s->d_pair_a, (size_t)np));
    A(dev_alloc(&s->d_pair_b, (size_t)np));
    A(dev_alloc(&s->d_state, (size_t)6 * ne * nr));
    A(dev_alloc(&s->d_scalars, (size_t)2 * ne));
    A(dev_alloc(&s->d_ints, (size_t)7 * ne));
    A(dev_alloc(&s->d_converged, (size_t)ne));
    A(dev_alloc(&s->d_touched, (size_t)ne * np));
    A(dev_alloc(&s->d_active, (size_t)ne));
    const int ng = batch->gset ? sdfc_host::ground_set_count(batch->gset) : 0;
    s->n_grounds = ng;
    A(dev_alloc(&s->d_touched_ground, (size_t)ne * ng));
    A(dev_alloc(&s->d_ground_body, (size_t)ng));
    A(dev_alloc(&s->d_fext, (size_t)ne * nb * 6));
    A(dev_alloc(&s->d_unfinished, (size_t)sdfc_sim::kMaxLanes));
    A(dev_alloc(&s->d_counters, (size_t)4));
    {   // SDFC_SKIN=0/1: temporal-coherence culling of the detection kernel (BatchDev::skin_*); results are identical
        const char* skin_env = std::getenv("SDFC_SKIN");
        if (skin_env ? std::atoi(skin_env) != 0 : (SDFC_SKIN_DEFAULT != 0)) A(sdfc_host::batch_enable_skin(batch));
    }
    if (st == SDFC_OK) {
        // SDFC_SIM_LANES overrides the lane count (1 = everything on the context's stream)
        const char* lanes_env = std::getenv("SDFC_SIM_LANES");
        s->n_lanes = lanes_env ? std::max(1, std::min((int)sdfc_sim::kMaxLanes, std::atoi(lanes_env))) : (ne >= 128 ? 2 : 1);
        s->n_lanes = std::min(s->n_lanes, std::max(1, ne / 32));
        A(sdfc_host::batch_ensure_plan_slots(batch, s->n_lanes));   // every lane plans its own force work lists
        cudaError_t e = cudaMallocHost((void**)&s->h_unfinished, sizeof(int) * sdfc_sim::kMaxLanes);
        for (int l = 0; l < s->n_lanes && e == cudaSuccess; ++l) {
            e = cudaStreamCreateWithFlags(&s->lane_stream[l], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->lane_event[l], cudaEventDisableTiming);
            for (int a = 0; a < 2 && e == cudaSuccess; ++a) e = cudaStreamCreateWithFlags(&s->lane_aux[l][a], cudaStreamNonBlocking);
            for (int a = 0; a < 4 && e == cudaSuccess; ++a) e = cudaEventCreateWithFlags(&s->lane_ev[l][a], cudaEventDisableTiming);
        }
        if (const char* g = std::getenv("SDFC_SIM_GRAPH")) s->use_graph = std::atoi(g) != 0;   // 0: launch every round kernel by kernel
        // CTA-per-environment variant of the round kernel: the round kernel is a chain of small dependent statements,
        // so while the teams of a lane's unfinished environments still fit the device side by side (4 CTAs per SM)
        // its duration is the latency of ONE environment and four warps per environment shorten it (measured on
        // 1024 rollouts in 2 lanes: 62.3 -> 58.2 ms translational, 66.3 -> 58.8 ms free3d-exp per forward(100));
        // eight warps per environment while one CTA per SM covers the lane (128 rollouts: 18.8 -> 17.9 ms; at 1024
        // they cost 45.2 -> 49.6 ms); larger lanes start on the warp-per-environment variant and switch when enough
        // environments are done.
        // SDFC_SIM_TAIL overrides the threshold (0 = never).
        s->tail_envs = 4 * sdfc::device_sm_count();
        if (const char* t = std::getenv("SDFC_SIM_TAIL")) s->tail_envs = std::max(0, std::atoi(t));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->fork_event, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_bodies, hb.data(), sizeof(BodyDev) * nb, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_pair_a, batch->body_a.data(), sizeof(int) * np, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_pair_b, batch->body_b.data(), sizeof(int) * np, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(s->d_state, 0, sizeof(double) * 6 * ne * nr);
        if (e == cudaSuccess) e = cudaMemset(s->d_scalars, 0, sizeof(double) * 2 * ne);
        if (e == cudaSuccess) e = cudaMemset(s->d_ints, 0, sizeof(int) * 7 * ne);
        if (e == cudaSuccess) e = cudaMemset(s->d_converged, 1, ne);
        if (e == cudaSuccess) e = cudaMemset(s->d_touched, 0, (size_t)ne * np);
        if (e == cudaSuccess && ng) e = cudaMemset(s->d_touched_ground, 0, (size_t)ne * ng);
        if (e == cudaSuccess && ng) e = cudaMemcpy(s->d_ground_body, sdfc_host::ground_set_bodies(batch->gset), sizeof(int) * ng, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(s->d_fext, 0, sizeof(double) * ne * nb * 6);
        if (e == cudaSuccess) e = cudaMemset(s->d_counters, 0, sizeof(unsigned long long) * 4);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();  // null-stream work done before the lane / context streams use it
        if (e != cudaSuccess) st = fail(SDFC_ERR_CUDA, std::string("sdfc_sim_create: ") + cudaGetErrorString(e));
    }
    if (st != SDFC_OK) { sdfc_sim_destroy(s); return st; }
    SimDev& D = s->dev;
    D.bodies = s->d_bodies; D.pair_a = s->d_pair_a; D.pair_b = s->d_pair_b;
    D.n_bodies = nb; D.n_pairs = np; D.n_envs = ne; D.nr = nr; D.n_dyn = nd;
    D.max_iter = std::max(20 * nr, 100);  // Simulation.cpp:1009
    D.h = h;
    for (int a = 0; a < 3; ++a) D.grav[a] = gravity[a];
    const size_t n = (size_t)ne * nr;
    D.q = s->d_state; D.qdot = s->d_state + n; D.q0 = s->d_state + 2 * n; D.qdot0 = s->d_state + 3 * n;
    D.x = s->d_state + 4 * n; D.dx = s->d_state + 5 * n;
    D.gnorm = s->d_scalars; D.alpha = s->d_scalars + ne;
    D.phase = s->d_ints; D.trial = s->d_ints + ne; D.iter = s->d_ints + 2 * ne; D.strike = s->d_ints + 3 * ne;
    D.steps_left = s->d_ints + 4 * ne; D.step_idx = s->d_ints + 5 * ne; D.env_mode = s->d_ints + 6 * ne;
    D.converged = s->d_converged; D.touched = s->d_touched; D.fext = s->d_fext; D.active = nullptr;
    D.E = batch->d_E; D.phi = batch->d_phi;
    D.fm_sub = batch->d_fm_sub; D.Km_sub = batch->d_Km_sub; D.Dm_sub = batch->d_Dm_sub; D.n_active = batch->d_n_active;
    D.n_grounds = ng; D.ground_body = s->d_ground_body; D.touched_ground = s->d_touched_ground;
    D.g_out = ng ? sdfc_host::ground_set_out(batch->gset) : nullptr;
    D.g_n_active = ng ? sdfc_host::ground_set_n_active(batch->gset) : nullptr;
    D.q_traj = nullptr; D.qdot_traj = nullptr;
    D.n_unfinished = s->d_unfinished; D.counters = s->d_counters;
    D.env0 = 0; D.env_count = ne;
    *out = s;
    return SDFC_OK;
}

sdfc_status sdfc_sim_destroy(sdfc_sim* s) {
    if (!s) return SDFC_OK;
    cudaSetDevice(s->ctx->device);
    cudaFree(s->d_bodies); cudaFree(s->d_pair_a); cudaFree(s->d_pair_b); cudaFree(s->d_state); cudaFree(s->d_scalars);
    cudaFree(s->d_ints); cudaFree(s->d_converged); cudaFree(s->d_touched); cudaFree(s->d_active); cudaFree(s->d_touched_ground); cudaFree(s->d_ground_body); cudaFree(s->d_fext); cudaFree(s->d_unfinished);
    cudaFree(s->d_counters); cudaFree(s->d_traj_q); cudaFree(s->d_traj_qd);
    if (s->h_unfinished) cudaFreeHost(s->h_unfinished);
    for (int l = 0; l < sdfc_sim::kMaxLanes; ++l) {
        for (int v = 0; v < 3; ++v) if (s->lane_graph[l][v]) cudaGraphExecDestroy(s->lane_graph[l][v]);
        if (s->lane_stream[l]) cudaStreamDestroy(s->lane_stream[l]);
        if (s->lane_event[l]) cudaEventDestroy(s->lane_event[l]);
        for (int a = 0; a < 2; ++a) if (s->lane_aux[l][a]) cudaStreamDestroy(s->lane_aux[l][a]);
        for (int a = 0; a < 4; ++a) if (s->lane_ev[l][a]) cudaEventDestroy(s->lane_ev[l][a]);
    }
    if (s->fork_event) cudaEventDestroy(s->fork_event);
    delete s;
    return SDFC_OK;
}

int sdfc_sim_ndof_r(const sdfc_sim* s) { return s ? s->nr : 0; }

sdfc_status sdfc_sim_set_state(sdfc_sim* s, const double* q_host, const double* qdot_host) {
    SDFC_REQUIRE(s && q_host && qdot_host, "sdfc_sim_set_state: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    const size_t bytes = sizeof(double) * (size_t)s->n_envs * s->nr;
    SDFC_CUDA(cudaMemcpyAsync(s->dev.q, q_host, bytes, cudaMemcpyHostToDevice, s->ctx->stream));
    SDFC_CUDA(cudaMemcpyAsync(s->dev.qdot, qdot_host, bytes, cudaMemcpyHostToDevice, s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_state(sdfc_sim* s, double* q_host, double* qdot_host) {
    SDFC_REQUIRE(s && (q_host || qdot_host), "sdfc_sim_get_state: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    const size_t bytes = sizeof(double) * (size_t)s->n_envs * s->nr;
    if (q_host) SDFC_CUDA(cudaMemcpyAsync(q_host, s->dev.q, bytes, cudaMemcpyDeviceToHost, s->ctx->stream));
    if (qdot_host) SDFC_CUDA(cudaMemcpyAsync(qdot_host, s->dev.qdot, bytes, cudaMemcpyDeviceToHost, s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_set_external_force(sdfc_sim* s, const double* f_host) {
    SDFC_REQUIRE(s && f_host, "sdfc_sim_set_external_force: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    SDFC_CUDA(cudaMemcpyAsync(s->d_fext, f_host, sizeof(double) * (size_t)s->n_envs * s->n_bodies * 6, cudaMemcpyHostToDevice,
                              s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_clear_contact_history(sdfc_sim* s) {
    SDFC_REQUIRE(s != nullptr, "sdfc_sim_clear_contact_history: sim is NULL");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    SDFC_CUDA(cudaMemsetAsync(s->d_touched, 0, (size_t)s->n_envs * s->n_pairs, s->ctx->stream));
    if (s->n_grounds)
        SDFC_CUDA(cudaMemsetAsync(s->d_touched_ground, 0, (size_t)s->n_envs * s->n_grounds, s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_set_active(sdfc_sim* s, const uint8_t* active_host) {
    SDFC_REQUIRE(s != nullptr, "sdfc_sim_set_active: sim is NULL");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    s->use_active = active_host != nullptr;
    if (active_host) {
        SDFC_CUDA(cudaMemcpyAsync(s->d_active, active_host, (size_t)s->n_envs, cudaMemcpyHostToDevice, s->ctx->stream));
        SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    }
    return SDFC_OK;
}

sdfc_status sdfc_sim_forward(sdfc_sim* s, int n_steps, double* q_traj_host, double* qdot_traj_host) {
    SDFC_REQUIRE(s != nullptr, "sdfc_sim_forward: sim is NULL");
    SDFC_REQUIRE(n_steps >= 0, "sdfc_sim_forward: negative step count");
    if (n_steps == 0) return SDFC_OK;
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    const size_t per_step = (size_t)s->n_envs * s->nr;
    const bool want_traj = q_traj_host || qdot_traj_host;
    if (want_traj && s->traj_cap < (size_t)n_steps) {
        cudaFree(s->d_traj_q); cudaFree(s->d_traj_qd);
        s->d_traj_q = s->d_traj_qd = nullptr; s->traj_cap = 0;
        sdfc_status a = dev_alloc(&s->d_traj_q, per_step * n_steps);
        if (a != SDFC_OK) return a;
        a = dev_alloc(&s->d_traj_qd, per_step * n_steps);
        if (a != SDFC_OK) return a;
        s->traj_cap = (size_t)n_steps;
    }
    {
        sdfc_status pr = sdfc_host::batch_prepare(s->batch);   // pair table / parameter changes since the last call
        if (pr != SDFC_OK) return pr;
    }
    SimDev D = s->dev;
    D.q_traj = q_traj_host ? s->d_traj_q : nullptr;
    D.qdot_traj = qdot_traj_host ? s->d_traj_qd : nullptr;
    D.active = s->use_active ? s->d_active : nullptr;
    D.B = s->batch->dev;
    D.B.env_mode = D.env_mode;
    D.dm = (s->batch->any_fric || s->batch->any_damp) ? 1 : 0;
    // everything a captured round depends on besides the (fixed) device buffers
    uint64_t graph_key = 1469598103934665603ull;
    {
        const uint64_t parts[8] = {(uint64_t)(uintptr_t)D.q_traj, (uint64_t)(uintptr_t)D.qdot_traj, (uint64_t)(uintptr_t)D.active,
                                   (uint64_t)s->batch->any_fric * 2 + (uint64_t)s->batch->any_damp, (uint64_t)(uintptr_t)D.B.skin_margin,
                                   (uint64_t)D.B.any_bricks, (uint64_t)(uintptr_t)s->ctx->stream, (uint64_t)D.B.brick_cap};
        for (uint64_t v : parts) { graph_key ^= v; graph_key *= 1099511628211ull; }
        if (graph_key == 0) graph_key = 1;
    }
    // Lanes: each environment range advances through its own rounds on its own stream; the
    // host feeds the lanes round-robin, so the 32-CTA assembly / solve kernels of one lane run
    // under the contact kernels of another. Results do not depend on the lane count.
    const int L = s->n_lanes;
    SDFC_CUDA(cudaMemsetAsync(s->d_unfinished, 0, sizeof(int) * sdfc_sim::kMaxLanes, st));
    SDFC_CUDA(cudaEventRecord(s->fork_event, st));
    int lane_env0[sdfc_sim::kMaxLanes], lane_cnt[sdfc_sim::kMaxLanes];
    bool lane_live[sdfc_sim::kMaxLanes];
    for (int l = 0; l < L; ++l) {
        lane_env0[l] = (int)((int64_t)s->n_envs * l / L);
        lane_cnt[l] = (int)((int64_t)s->n_envs * (l + 1) / L) - lane_env0[l];
        lane_live[l] = lane_cnt[l] > 0;
        if (!lane_live[l]) continue;
        cudaStream_t ls = (L == 1) ? st : s->lane_stream[l];
        if (L > 1) SDFC_CUDA(cudaStreamWaitEvent(ls, s->fork_event, 0));
        SimDev Dl = D;
        Dl.env0 = lane_env0[l]; Dl.env_count = lane_cnt[l]; Dl.n_unfinished = s->d_unfinished + l;
        k_sim_begin<<<(lane_cnt[l] + 31) / 32, 32, 0, ls>>>(Dl, n_steps);
        s->ctx->launches += 1;
    }
    // every step needs at least two rounds (one Newton evaluation, one line-search
    // trial), so the unfinished-counter is only polled every second round
    const int64_t guard = (int64_t)n_steps * (D.max_iter * (kMaxLS + 1) + 1);
    int64_t depth = 0;
    sdfc_status rc = SDFC_OK;
    // The unfinished-counter of a lane is read back with a lag: the copy enqueued after round r
    // is only waited for before round r + 2 is issued, so the device always has queued work while
    // the host looks at an older answer (a finished lane runs at most two empty rounds).
    bool lane_pending[sdfc_sim::kMaxLanes] = {false, false, false, false};
    int lane_tail[sdfc_sim::kMaxLanes];   // round-kernel variant of the lane (sdfc_sim::team_variant)
    for (int l = 0; l < sdfc_sim::kMaxLanes; ++l) lane_tail[l] = l < L ? s->team_variant(lane_cnt[l]) : 0;
    int64_t lane_round[sdfc_sim::kMaxLanes] = {0, 0, 0, 0};
    for (int64_t round = 0; round < guard && rc == SDFC_OK; ++round) {
        bool any = false;
        for (int l = 0; l < L && rc == SDFC_OK; ++l) {
            if (!lane_live[l]) continue;
            cudaStream_t ls = (L == 1) ? st : s->lane_stream[l];
            if ((round & 1) && lane_pending[l]) {
                if (cudaEventSynchronize(s->lane_event[l]) != cudaSuccess) { rc = fail(SDFC_ERR_CUDA, "sdfc_sim_forward: lane sync failed"); break; }
                lane_pending[l] = false;
                if (s->h_unfinished[l] <= 0) { lane_live[l] = false; continue; }
                lane_tail[l] = s->team_variant(s->h_unfinished[l]);
            }
            any = true;
            SimDev Dl = D;
            Dl.env0 = lane_env0[l]; Dl.env_count = lane_cnt[l]; Dl.n_unfinished = s->d_unfinished + l;
            Dl.B.plan[0] = s->batch->plan_slot[l][0]; Dl.B.plan[1] = s->batch->plan_slot[l][1];
            // One round = k_sim_round (state machine on the evaluation of the previous round - none in the first one -
            // then poses, evaluation modes, zeroed counters and culling decisions of the next candidate) followed by
            // the contact kernels that evaluate it. The steady-state round is captured once into a CUDA graph
            // (fork / join: value and derivative force kernels and the ground kernels side by side) and replayed.
            auto issue = [&](int kind) -> sdfc_status {
                sdfc_status r = launch_round(s, Dl, kind, ls, lane_tail[l]);
                if (r != SDFC_OK) return r;
                return sdfc_host::batch_run_stepper(s->batch, D.E, D.phi, D.env_mode, lane_env0[l], lane_cnt[l], l, ls, s->lane_aux[l][0],
                                                    s->lane_aux[l][1], s->lane_ev[l]);
            };
            if (lane_round[l] == 0 || !s->use_graph) {
                rc = issue(lane_round[l] == 0 ? ROUND_FIRST : ROUND_FULL);
            } else {
                const int v = lane_tail[l];
                if (s->lane_graph[l][v] == nullptr || s->lane_graph_key[l][v] != graph_key) {
                    if (s->lane_graph[l][v]) { cudaGraphExecDestroy(s->lane_graph[l][v]); s->lane_graph[l][v] = nullptr; }
                    cudaGraph_t g = nullptr;
                    const int64_t launches0 = s->ctx->launches;
                    if (cudaStreamBeginCapture(ls, cudaStreamCaptureModeRelaxed) != cudaSuccess) { rc = fail(SDFC_ERR_CUDA, "sdfc_sim_forward: graph capture failed to start"); break; }
                    rc = issue(ROUND_FULL);
                    const cudaError_t ce = cudaStreamEndCapture(ls, &g);
                    s->ctx->launches = launches0;   // captured, not launched
                    if (rc == SDFC_OK && (ce != cudaSuccess || g == nullptr)) rc = fail(SDFC_ERR_CUDA, std::string("sdfc_sim_forward: graph capture: ") + cudaGetErrorString(ce));
                    if (rc == SDFC_OK && cudaGraphInstantiate(&s->lane_graph[l][v], g, 0) != cudaSuccess) rc = fail(SDFC_ERR_CUDA, "sdfc_sim_forward: graph instantiation failed");
                    if (g) cudaGraphDestroy(g);
                    if (rc != SDFC_OK) break;
                    s->lane_graph_key[l][v] = graph_key;
                }
                if (cudaGraphLaunch(s->lane_graph[l][v], ls) != cudaSuccess) { rc = fail(SDFC_ERR_CUDA, "sdfc_sim_forward: graph launch failed"); break; }
                s->ctx->launches += 5 + (s->n_grounds ? 2 : 0);
            }
            if (rc != SDFC_OK) break;
            lane_round[l] += 1;
            if (round & 1) {
                if (cudaMemcpyAsync(s->h_unfinished + l, s->d_unfinished + l, sizeof(int), cudaMemcpyDeviceToHost, ls) != cudaSuccess ||
                    cudaEventRecord(s->lane_event[l], ls) != cudaSuccess)
                    rc = fail(SDFC_ERR_CUDA, "sdfc_sim_forward: poll failed");
                lane_pending[l] = true;
            }
        }
        if (!any) break;
        depth += 1;
    }
    s->rounds += depth;
    if (rc != SDFC_OK) { cudaDeviceSynchronize(); return rc; }
    if (L > 1)
        for (int l = 0; l < L; ++l) {  // join: the context stream continues after every lane
            SDFC_CUDA(cudaEventRecord(s->lane_event[l], s->lane_stream[l]));
            SDFC_CUDA(cudaStreamWaitEvent(st, s->lane_event[l], 0));
        }
    SDFC_CUDA(cudaGetLastError());
    if (q_traj_host) SDFC_CUDA(cudaMemcpyAsync(q_traj_host, s->d_traj_q, sizeof(double) * per_step * n_steps, cudaMemcpyDeviceToHost, st));
    if (qdot_traj_host) SDFC_CUDA(cudaMemcpyAsync(qdot_traj_host, s->d_traj_qd, sizeof(double) * per_step * n_steps, cudaMemcpyDeviceToHost, st));
    SDFC_CUDA(cudaStreamSynchronize(st));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_status(sdfc_sim* s, uint8_t* converged, uint8_t* touched) {
    SDFC_REQUIRE(s != nullptr, "sdfc_sim_get_status: sim is NULL");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    if (converged) SDFC_CUDA(cudaMemcpyAsync(converged, s->d_converged, s->n_envs, cudaMemcpyDeviceToHost, s->ctx->stream));
    if (touched) SDFC_CUDA(cudaMemcpyAsync(touched, s->d_touched, (size_t)s->n_envs * s->n_pairs, cudaMemcpyDeviceToHost, s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_ground_history(sdfc_sim* s, uint8_t* touched_ground) {
    SDFC_REQUIRE(s && touched_ground, "sdfc_sim_get_ground_history: NULL argument");
    SDFC_REQUIRE(s->n_grounds > 0, "sdfc_sim_get_ground_history: the simulation has no ground contacts");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    SDFC_CUDA(cudaMemcpyAsync(touched_ground, s->d_touched_ground, (size_t)s->n_envs * s->n_grounds, cudaMemcpyDeviceToHost,
                              s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_poses(sdfc_sim* s, double* E_host) {
    SDFC_REQUIRE(s && E_host, "sdfc_sim_get_poses: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    sdfc_status rc = launch_round(s, s->dev, ROUND_POSES, s->ctx->stream);
    if (rc != SDFC_OK) return rc;
    SDFC_CUDA(cudaMemcpyAsync(E_host, s->dev.E, sizeof(double) * (size_t)s->n_envs * s->n_bodies * 16, cudaMemcpyDeviceToHost,
                              s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return SDFC_OK;
}

sdfc_status sdfc_sim_get_stats(sdfc_sim* s, int64_t stats[4]) {
    SDFC_REQUIRE(s && stats, "sdfc_sim_get_stats: NULL argument");
    SDFC_CUDA(cudaSetDevice(s->ctx->device));
    unsigned long long c[4] = {0, 0, 0, 0};
    SDFC_CUDA(cudaMemcpyAsync(c, s->d_counters, sizeof(c), cudaMemcpyDeviceToHost, s->ctx->stream));
    SDFC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    stats[0] = s->rounds; stats[1] = (int64_t)c[0]; stats[2] = (int64_t)c[1]; stats[3] = (int64_t)c[2];
    return SDFC_OK;
}

}  // extern "C"
