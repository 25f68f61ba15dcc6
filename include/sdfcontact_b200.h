/* sdfcontact_b200.h — C-ABI of the B200 (sm_100a) SDF-contact hot path.
 *
 * This is the drop-in boundary for ONE path of the RedMax simulator inside
 * yunshengtian/Assemble-Them-All: point-vs-SDF contact detection and penalty
 * contact force / Jacobian evaluation. Plain pointers and sizes,
 * no C++/torch types. Every entry point cites the reference interface it
 * replaces (paths relative to simulation/redmax/ of the reference).
 *
 * Conventions (Eigen-native, so a C++ host passes `.data()` directly):
 *   - 4x4 poses E are COLUMN-MAJOR double[16]  (Body::_E_0i, Body/Body.h:52)
 *   - spatial velocities phi are double[6] = (omega, v) in body frame (Body.h:58)
 *   - fm_sub is double[12] = (contact body 6, SDF body 6); Km_sub/Dm_sub are
 *     COLUMN-MAJOR double[144]                  (ForceGeneralSDFContact.cpp:84-99)
 *   - SDF grids are linear, x fastest: a[i + ni*(j + nj*k)]
 *                                               (externals/sdfgen/array3.h:70-80)
 *   - contact points are AoS xyz in the contact body's frame
 *                                               (Body::_contact_points, Body.h:36)
 * All functions return SDFC_OK (0) or a negative sdfc_status; the message of
 * the last failure on the calling thread is available from sdfc_last_error().
 * There is NO CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef SDFCONTACT_B200_H
#define SDFCONTACT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SDFC_API
#else
#define SDFC_API __attribute__((visibility("default")))
#endif

typedef int sdfc_status;
#define SDFC_OK 0
#define SDFC_ERR_INVALID (-1)  /* bad argument */
#define SDFC_ERR_CUDA (-2)     /* CUDA runtime error (message has the detail) */
#define SDFC_ERR_NO_DEVICE (-3)

typedef struct sdfc_context sdfc_context; /* one device + stream + scratch */
typedef struct sdfc_sdf sdfc_sdf;         /* device copy of one body's voxel SDF */
typedef struct sdfc_points sdfc_points;   /* device copy of one body's contact points */
typedef struct sdfc_pair sdfc_pair;       /* one ForceGeneralSDFContact */
typedef struct sdfc_batch sdfc_batch;     /* E environments x P pairs, device-resident */
typedef struct sdfc_ground sdfc_ground;   /* one ForceGroundContact */

/* ---- library ---------------------------------------------------------- */
SDFC_API int sdfc_version(void);
/* Replaces throw_error()/`throw std::string` (Common.h:58-61): the host
 * rethrows in the reference's convention using this message. */
SDFC_API const char* sdfc_last_error(void);
SDFC_API int sdfc_device_count(void);

/* Page-locked host memory for the host-buffer entries: buffers allocated here (or
 * registered with cudaHostRegister / torch pin_memory) are DMA'd directly by
 * sdfc_batch_eval; pageable buffers go through an internal pinned staging copy. */
SDFC_API sdfc_status sdfc_pinned_alloc(size_t bytes, void** out);
SDFC_API sdfc_status sdfc_pinned_free(void* p);

/* ---- context ---------------------------------------------------------- */
SDFC_API sdfc_status sdfc_context_create(int device, sdfc_context** out);
SDFC_API sdfc_status sdfc_context_destroy(sdfc_context* ctx);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the
 * context's own stream. */
SDFC_API sdfc_status sdfc_context_set_stream(sdfc_context* ctx, void* cuda_stream);
SDFC_API sdfc_status sdfc_context_synchronize(sdfc_context* ctx);
/* Number of kernels this context has launched so far (bench bookkeeping). */
SDFC_API int64_t sdfc_context_launch_count(const sdfc_context* ctx);
/* Per-kernel device timing with CUDA events recorded on the launching stream
 * around each kernel of the batched path. Slots: 0 = detection (LOOP C),
 * 1 = force / Jacobian (LOOP D), 2 = split reduction, 3 = system scatter.
 * read_timing synchronises the stream, adds the elapsed times of every launch
 * since the previous read into ms[4] / counts[4], and optionally resets. */
SDFC_API sdfc_status sdfc_context_enable_timing(sdfc_context* ctx, int enable);
SDFC_API sdfc_status sdfc_context_read_timing(sdfc_context* ctx, double ms[4], int64_t counts[4], int reset);

/* Device self-test: n pseudo-random numerators divided by b with the library's
 * reciprocal-based exact division, compared bit for bit with the IEEE division.
 * mismatches must come back 0 (the cell index and the interpolation weights of
 * query_SDF, BodySDFObj.cpp:132-134, depend on it). */
/* Roofline denominators measured on this device (MEASURED_PEAKS.json covers HBM and bf16 only):
 * out[0] = FP64 DFMA throughput in GFLOP/s (8 independent chains per thread, all SMs),
 * out[1] = L2 -> SM read bandwidth in GB/s (every CTA streams a 48 MiB buffer with 128-bit loads),
 * out[2..3] reserved (0). Takes a few milliseconds. */
SDFC_API sdfc_status sdfc_measure_peaks(sdfc_context* ctx, double out[4]);
SDFC_API sdfc_status sdfc_selftest_exact_div(sdfc_context* ctx, int64_t n, uint64_t seed, double b, int64_t* mismatches);

/* ---- SDF grid: fields of BodySDFObj (Body/BodySDFObj.h:13-16) ---------- */
/* grid_host: ni*nj*nk doubles. If every value is exactly representable in
 * fp32 (true for freshly generated grids, BodySDFObj.cpp:81-83) an fp32 copy
 * is stored as well and used for gathers (values are widened back exactly);
 * grids reloaded from the text cache (BodySDFObj.cpp:86-103) stay fp64. */
SDFC_API sdfc_status sdfc_sdf_create(sdfc_context* ctx, int ni, int nj, int nk,
                                     const double min_box[3], const double max_box[3],
                                     double dx, double dy, double dz, double col_th,
                                     const double* grid_host, sdfc_sdf** out);
SDFC_API sdfc_status sdfc_sdf_set_col_th(sdfc_sdf* sdf, double col_th);
/* 1 if the fp32-lossless copy is in use, else 0. */
SDFC_API int sdfc_sdf_is_fp32(const sdfc_sdf* sdf);
/* device storage of the grid: 0 = fp64 only, 1 = fp32 cell records (32 B per cell), 2 = fp32 bricks (4 B per
 * node, 4x4x4 bricks staged per detection tile with TMA), 3 = both (grids whose records exceed ~96 MB and fit
 * 6 GiB: bricks for detection, records for per-point gathers); the environment variable
 * SDFC_GRID_LAYOUT=records|bricks|both overrides the size rule */
SDFC_API int sdfc_sdf_layout_kind(const sdfc_sdf* sdf);
SDFC_API sdfc_status sdfc_sdf_destroy(sdfc_sdf* sdf);

/* Batched raw sampling = SDFMesh::distance / min_distance (Mesh.cpp:143-198,
 * python_interface.cpp:283-292): no col_th, FLT_MAX outside the box.
 * X_host: n x 3 AoS in the SDF's own frame. Host buffers in and out. */
SDFC_API sdfc_status sdfc_sdf_distance(sdfc_sdf* sdf, int64_t n, const double* X_host, double* out_host);
SDFC_API sdfc_status sdfc_sdf_min_distance(sdfc_sdf* sdf, int64_t n, const double* X_host, double* out_min);
/* Device-resident form: SoA inputs already in HBM, result left in HBM. */
/* BodySDFObj::query_SDF(x, true), query_dSDF and query_ddSDF (Body/BodySDFObj.cpp:123-205) at n points
 * (AoS, grid frame): sdf_acc_out[n], grad_out[3n], hess_out[9n] row-major; any output may be NULL.
 * fast != 0 evaluates the shared-stencil path of the planner kernel, 0 the sample-by-sample path;
 * both are bit-identical to the reference (rows a6 / a7 of the scope table). */
SDFC_API sdfc_status sdfc_sdf_stencil(sdfc_sdf* sdf, int64_t n, const double* X_host, int fast, double* sdf_acc_out,
                                      double* grad_out, double* hess_out);
SDFC_API sdfc_status sdfc_sdf_distance_device(sdfc_sdf* sdf, int64_t n, const double* dx, const double* dy,
                                              const double* dz, double* dout);

/* ---- SDF construction: sdfgen::make_level_set3 + BodySDFObj::precompute_SDF ---- */
/* Drop-in for sdfgen::make_level_set3 (externals/sdfgen/makelevelset3.h:15-17, .cpp:164-226)
 * as called by BodySDFObj::precompute_SDF (Body/BodySDFObj.cpp:75) and SDFMesh
 * (Mesh.cpp:54-100). tri: [n_tri][3] vertex indices; x: [n_vert][3] fp32 vertices;
 * phi_out: ni*nj*nk fp32, x fastest (Array3f). The result is bit-identical to the CPU
 * routine (exact band, 2 x 8 sweeps in the reference's dependency order, parity signs).
 * Host buffers in and out. */
SDFC_API sdfc_status sdfc_make_level_set3(sdfc_context* ctx, int n_tri, const uint32_t* tri, int n_vert, const float* x,
                                          const float origin[3], float dx, float dy, float dz, int ni, int nj, int nk,
                                          int exact_band, float* phi_out);
/* Grid layout of BodySDFObj::precompute_SDF (BodySDFObj.cpp:41-70): V = body-frame
 * vertices [n_vert][3] fp64, dx = requested spacing, res = min cells per axis (20,
 * Simulation_Constructor.cpp:568). Outputs: spacing[3] = min(dx, bbox/res) per axis,
 * fp32 min_box / max_box with one cell of padding, dims[3]. Host only. */
SDFC_API sdfc_status sdfc_sdf_layout(int n_vert, const double* V, double dx, int res, double spacing[3], float min_box[3],
                                     float max_box[3], int32_t dims[3]);
/* The `.sdf` text cache of BodySDFObj::save_SDF / load_SDF (BodySDFObj.cpp:86-116): same
 * token order and the same 6-significant-digit formatting (so a reloaded grid is NOT
 * fp32-lossless and stays fp64 in sdfc_sdf_create). Host only. */
SDFC_API sdfc_status sdfc_sdf_save_text(const char* path, int ni, int nj, int nk, const double min_box[3],
                                        const double max_box[3], double dx, double dy, double dz, const double* grid);
SDFC_API sdfc_status sdfc_sdf_load_text_header(const char* path, int32_t dims[3]);
SDFC_API sdfc_status sdfc_sdf_load_text(const char* path, double min_box[3], double max_box[3], double spacing[3],
                                        double* grid, int64_t capacity);

/* ---- contact points: Body::_contact_points (Body/Body.h:36,129) -------- */
/* morton_sort != 0 stores a spatially sorted SoA copy for the force path and batched detection (the
 * original order is kept for every per-point output): 1 = 30-bit Morton order, 2 = k-d order whose
 * 2048-point detection tiles and 32-point words are compact boxes (what the TMA brick staging wants). */
SDFC_API sdfc_status sdfc_points_create(sdfc_context* ctx, int n, const double* xyz_host, int morton_sort,
                                        sdfc_points** out);
SDFC_API int sdfc_points_count(const sdfc_points* pts);
SDFC_API sdfc_status sdfc_points_destroy(sdfc_points* pts);

/* ---- pair: ForceGeneralSDFContact (Force/ForceGeneralSDFContact.h:13-31) */
SDFC_API sdfc_status sdfc_pair_create(sdfc_context* ctx, sdfc_points* contact_points, sdfc_sdf* sdf,
                                      double kn, double kt, double mu, double damping, sdfc_pair** out);
SDFC_API sdfc_status sdfc_pair_set_stiffness(sdfc_pair* pair, double kn, double kt); /* .h:28 */
SDFC_API sdfc_status sdfc_pair_set_friction(sdfc_pair* pair, double mu);             /* .h:29 */
SDFC_API sdfc_status sdfc_pair_set_damping(sdfc_pair* pair, double damping);         /* .h:30 */
SDFC_API sdfc_status sdfc_pair_set_scale(sdfc_pair* pair, double scale);             /* .h:31 */
/* Body::_contact_scale of the contact body (Body/Body.h:47, Robot.cpp:462-472). */
SDFC_API sdfc_status sdfc_pair_set_contact_scale(sdfc_pair* pair, double contact_scale);
SDFC_API sdfc_status sdfc_pair_destroy(sdfc_pair* pair);

/* collision_detection_general_SDF (CollisionDetection/CollisionDetection.cpp:124-147)
 * for every point, in point order: d_out[n] = SDF_body->distance(xw),
 * cell_out[3n] = query_grid_location, active_out[n] = (d < 0). Any of the three
 * output pointers may be NULL. n_active receives the contact count. Host buffers. */
SDFC_API sdfc_status sdfc_pair_query(sdfc_pair* pair, const double E1[16], const double E2[16],
                                     double* d_out, int32_t* cell_out, uint8_t* active_out, int32_t* n_active);

/* ForceGeneralSDFContact::computeForce (ForceGeneralSDFContact.cpp:46-67,147-219):
 * detection + value-only force. fm_sub[12] is OVERWRITTEN with the pair's
 * (contact body, SDF body) blocks; the caller scatters/accumulates (see
 * sdfc_scatter_*). n_active > 0 means both bodies record contact history
 * (CollisionDetection.cpp:143-146). */
SDFC_API sdfc_status sdfc_pair_force(sdfc_pair* pair, const double E1[16], const double phi1[6],
                                     const double E2[16], const double phi2[6],
                                     double fm_sub[12], int32_t* n_active);

/* ForceGeneralSDFContact::computeForceWithDerivative (.cpp:69-105,221-447). */
SDFC_API sdfc_status sdfc_pair_force_deriv(sdfc_pair* pair, const double E1[16], const double phi1[6],
                                           const double E2[16], const double phi2[6],
                                           double fm_sub[12], double Km_sub[144], double Dm_sub[144],
                                           int32_t* n_active);

/* The accumulate (+=) half of the public wrappers (.cpp:60-61,88-99): host-side
 * helper so a Force plug-in keeps the reference's semantics. System arrays are
 * column-major ndof_m x ndof_m. */
SDFC_API void sdfc_scatter_force(int idx_contact, int idx_sdf, const double fm_sub[12], double* fm_sys);
SDFC_API void sdfc_scatter_deriv(int ndof_m, int idx_contact, int idx_sdf, const double Km_sub[144],
                                 const double Dm_sub[144], double* Km_sys, double* Dm_sys);

/* ---- ground: ForceGroundContact (Force/ForceGroundContact.h, .cpp:9-38) ---- */
/* The second Force plug-in of the free-fall configuration (examples/test_joint_sim.py:
 * 36,54): contact points of one body against the plane through E_g's origin with normal
 * E_g's z axis. E_g: column-major double[16] (Simulation_Constructor.cpp:172-190). */
SDFC_API sdfc_status sdfc_ground_create(sdfc_context* ctx, sdfc_points* contact_points, const double E_g[16],
                                        double kn, double kt, double mu, double damping, sdfc_ground** out);
SDFC_API sdfc_status sdfc_ground_set_transform(sdfc_ground* g, const double E_g[16]);   /* .cpp:22 */
SDFC_API sdfc_status sdfc_ground_set_stiffness(sdfc_ground* g, double kn, double kt);   /* .cpp:26 */
SDFC_API sdfc_status sdfc_ground_set_friction(sdfc_ground* g, double mu);               /* .cpp:31 */
SDFC_API sdfc_status sdfc_ground_set_damping(sdfc_ground* g, double damping);           /* .cpp:35 */
SDFC_API sdfc_status sdfc_ground_destroy(sdfc_ground* g);
/* ForceGroundContact::computeForce (.cpp:39-57,106-153): detection
 * (collision_detection_ground_object, CollisionDetection.cpp:37-68: active iff d <= 0)
 * + force. fm_sub[6] is OVERWRITTEN with the body's block; n_active > 0 means the body
 * records "ground" in its contact history. */
SDFC_API sdfc_status sdfc_ground_force(sdfc_ground* g, const double E[16], const double phi[6], double fm_sub[6],
                                       int32_t* n_active);
/* ForceGroundContact::computeForceWithDerivative (.cpp:59-81,155-242): Km_sub / Dm_sub are
 * the body-diagonal 6x6 blocks, COLUMN-MAJOR double[36]. */
SDFC_API sdfc_status sdfc_ground_force_deriv(sdfc_ground* g, const double E[16], const double phi[6], double fm_sub[6],
                                             double Km_sub[36], double Dm_sub[36], int32_t* n_active);

/* ---- batched: E independent environments, device-resident -------------- */
/* An environment is one Simulation's worth of bodies: n_bodies poses, and
 * n_pairs contact forces pair[p] = (body_a[p] -> contact points, body_b[p] ->
 * SDF). All environments share geometry; only poses/velocities differ. This is
 * Robot::computeForce[WithDerivative]'s loop over _forces (Robot.cpp:513-551)
 * for E simulations at once. */
SDFC_API sdfc_status sdfc_batch_create(sdfc_context* ctx, int n_envs, int n_bodies, int n_pairs,
                                       sdfc_pair* const* pairs, const int32_t* body_a, const int32_t* body_b,
                                       sdfc_batch** out);
SDFC_API sdfc_status sdfc_batch_destroy(sdfc_batch* batch);
/* Adds the environment's ForceGroundContact forces (ground g acts on body[g]); they are
 * evaluated by every sdfc_batch_eval* call and by the stepper, and summed into the system
 * arrays together with the SDF pairs (Robot.cpp:513-551 loops every force). Call once,
 * before the first evaluation; the batch does not own the ground objects. */
SDFC_API sdfc_status sdfc_batch_set_grounds(sdfc_batch* batch, int n_grounds, sdfc_ground* const* grounds,
                                            const int32_t* body);
/* Running totals of the stepper's temporal-coherence culling in the detection kernel (all 0 when
 * it is off; SDFC_SKIN=0/1 at sdfc_sim_create): stats[0] = 32-point words skipped because
 * Lipschitz x displacement stayed below their measured margin, stats[1] = words whose margin was
 * used up (evaluated in full), stats[2] = (environment, pair) margin rebuilds, stats[3] = (environment,
 * pair) evaluations skipped as a whole (smallest margin of the pair against the displacement bound of
 * the whole point set). The cull never changes a result: a skipped word has no point with d < 0
 * (CollisionDetection.cpp:137-138). */
SDFC_API sdfc_status sdfc_batch_get_skin_stats(sdfc_batch* batch, int64_t stats[4]);
/* Ground contact counts of the last evaluation: [n_envs][n_grounds], host. */
/* Running totals of the TMA-staged detection (bricked grids): detection tiles whose voxel brick was staged in
 * shared memory, tiles whose brick exceeded the budget (global gathers), tiles outside the grid box. */
SDFC_API sdfc_status sdfc_batch_get_brick_stats(sdfc_batch* batch, int64_t stats[3]);
SDFC_API sdfc_status sdfc_batch_get_ground_active(sdfc_batch* batch, int32_t* n_active_out);

/* Host-buffer entry (the plug-in-facing call): E_host [n_envs][n_bodies][16],
 * phi_host [n_envs][n_bodies][6]. Outputs, per environment, the caller-zeroed
 * -then-accumulated system arrays exactly as Robot hands them to the forces:
 * fm_out [n_envs][6*n_bodies]; Km_out, Dm_out [n_envs][(6*n_bodies)^2] column-
 * major (NULL when deriv == 0); n_active_out [n_envs][n_pairs] or NULL. Arrays are
 * overwritten (contact contribution only). Includes H2D and D2H copies; Dm is not
 * transferred when it is identically zero (no friction, damping or ground force). */
SDFC_API sdfc_status sdfc_batch_eval(sdfc_batch* batch, const double* E_host, const double* phi_host, int deriv,
                                     double* fm_out, double* Km_out, double* Dm_out, int32_t* n_active_out);

/* Device-resident entry: same arrays, all in HBM; asynchronous on the context
 * stream. Dm_dev may be NULL (skipped). */
/* Same evaluation, results PER FORCE instead of accumulated: fm_sub_out[env][pair][12], Km_sub_out / Dm_sub_out
 * [env][pair][144] (column-major 12 x 12, deriv only), n_active_out[env][pair]; for the batch's ground contacts
 * ground_out[env][ground][78] = fm 6 | Km 36 | Dm 36 (column-major 6 x 6) and ground_active_out[env][ground].
 * One upload, one synchronisation. This is what a host whose plug-in interface is per force
 * (redmax::Force::computeForce[WithDerivative], Force/Force.h:9-58) calls once per Robot evaluation: every
 * force of the Robot then adds its own blocks (host/B200RobotBatch.h). */
SDFC_API sdfc_status sdfc_batch_eval_blocks(sdfc_batch* batch, const double* E_host, const double* phi_host, int deriv,
                                            double* fm_sub_out, double* Km_sub_out, double* Dm_sub_out,
                                            int32_t* n_active_out, double* ground_out, int32_t* ground_active_out);
SDFC_API sdfc_status sdfc_batch_eval_device(sdfc_batch* batch, const double* E_dev, const double* phi_dev,
                                            int deriv, double* fm_dev, double* Km_dev, double* Dm_dev,
                                            int32_t* n_active_dev);

/* Detection only over the whole batch (LOOP C): fills n_active_dev
 * [n_envs][n_pairs]; per-point outputs are not materialised. */
SDFC_API sdfc_status sdfc_batch_detect_device(sdfc_batch* batch, const double* E_dev, int32_t* n_active_dev);

/* Total detection queries one batch evaluation performs = n_envs * sum_p n_p. */
/* The 1-bit active mask of the last detection (batched kernel) for one (environment, pair), expanded
 * to active_out[n] (0 / 1) in the ORIGINAL contact-point order: the d < 0 test of
 * collision_detection_general_SDF (CollisionDetection.cpp:137-138). Synchronises. */
SDFC_API sdfc_status sdfc_batch_get_active(sdfc_batch* batch, int env, int pair, uint8_t* active_out);
SDFC_API int64_t sdfc_batch_queries_per_eval(const sdfc_batch* batch);

/* ---- batched BDF1 stepper: Simulation::forward for E environments -------- */
/* The caller side of the contact path, device-resident:
 * joint + body kinematics (Joint/Joint.cpp:95-141, JointFixed / JointTranslational /
 * JointFree3DExp + JointSphericalExp, Body/Body.cpp:138-181), body forces
 * (Body/Body.cpp:250-330), reduced assembly M, f, K, D (Simulation.cpp:379-503),
 * BDF1 residual and Jacobian (Simulation.cpp:1081-1112), Newton with the
 * reference's halving line search and strike rules, partial-pivot LU
 * (Simulation.cpp:999-1079), reparam (JointSphericalExp.cpp:237-264).
 * Robots are "flat": every link is a root link with its own joint, which is what
 * examples/run_joint_plan.py:140-178 and test_joint_sim.py:24-63 emit. Reduced
 * coordinates are ordered body by body (the reference's DFS joint order,
 * Robot.cpp:48-101); at most 2 non-fixed bodies / 12 reduced dofs. */
typedef struct sdfc_sim sdfc_sim;
#define SDFC_JOINT_FIXED 0         /* Joint/JointFixed.cpp */
#define SDFC_JOINT_TRANSLATIONAL 1 /* Joint/JointTranslational.cpp, q = p */
#define SDFC_JOINT_FREE3D_EXP 2    /* Joint/JointFree3DExp.cpp, q = (p, r) */

typedef struct sdfc_body_desc {
    int32_t joint_type;
    int32_t reserved;
    double E_pj_0[16]; /* Joint::_E_pj_0 (Joint/Joint.cpp:23-31), column-major */
    double E_ji[16];   /* Body::_E_ji (Body/Body.h), column-major */
    double inertia[6]; /* Body::_Inertia: principal moments (3), mass (3) */
    double mass;       /* Body::_mass */
} sdfc_body_desc;

/* `batch` supplies n_envs, n_bodies and the contact forces in force order
 * (Robot::_forces); bodies[n_bodies]; gravity = Options::_gravity; h = _h. */
SDFC_API sdfc_status sdfc_sim_create(sdfc_batch* batch, const sdfc_body_desc* bodies, const double gravity[3],
                                     double h, sdfc_sim** out);
SDFC_API sdfc_status sdfc_sim_destroy(sdfc_sim* sim);
SDFC_API int sdfc_sim_ndof_r(const sdfc_sim* sim);
/* Simulation::set_state (Simulation.cpp:73): q, qdot [n_envs][ndof_r], host. */
SDFC_API sdfc_status sdfc_sim_set_state(sdfc_sim* sim, const double* q_host, const double* qdot_host);
SDFC_API sdfc_status sdfc_sim_get_state(sdfc_sim* sim, double* q_host, double* qdot_host);
/* Simulation::set_body_external_force (Simulation.cpp:272): [n_envs][n_bodies][6]. */
SDFC_API sdfc_status sdfc_sim_set_external_force(sdfc_sim* sim, const double* f_host);
/* Simulation::reset's clear_all_contact_bodies (Simulation.cpp:896). */
SDFC_API sdfc_status sdfc_sim_clear_contact_history(sdfc_sim* sim);
/* Restrict the following sdfc_sim_forward calls to the environments with
 * active_host[env] != 0 (n_envs bytes); the others keep q, qdot, convergence flag
 * and contact history untouched and cost no contact work. NULL = every
 * environment steps. A batched BFS stage uses this to retire rollouts that hit
 * their stop rule (run_joint_plan.py:556-567) while the rest continue. */
SDFC_API sdfc_status sdfc_sim_set_active(sdfc_sim* sim, const uint8_t* active_host);
/* Simulation::forward(n_steps) (Simulation.cpp:914) for every environment.
 * q_traj_host / qdot_traj_host: [n_steps][n_envs][ndof_r] or NULL; with NULL the
 * call moves no per-step data to the host (state stays in HBM). */
SDFC_API sdfc_status sdfc_sim_forward(sdfc_sim* sim, int n_steps, double* q_traj_host, double* qdot_traj_host);
/* converged[n_envs]: Simulation::is_converged of the last step (Simulation.cpp:995);
 * touched[n_envs][n_pairs]: contact history since the last clear
 * (CollisionDetection.cpp:143-146). Either may be NULL. */
SDFC_API sdfc_status sdfc_sim_get_status(sdfc_sim* sim, uint8_t* converged, uint8_t* touched);
/* touched_ground[n_envs][n_grounds]: "ground" entries of the contact history
 * (CollisionDetection.cpp:67) since the last clear. */
SDFC_API sdfc_status sdfc_sim_get_ground_history(sdfc_sim* sim, uint8_t* touched_ground);
/* Simulation::get_body_E0i (Simulation.cpp:248) at the current state:
 * [n_envs][n_bodies][16] column-major. */
SDFC_API sdfc_status sdfc_sim_get_poses(sdfc_sim* sim, double* E_host);
/* Counters since creation: [0] evaluation rounds launched, [1] environment
 * evaluations with derivatives, [2] value-only environment evaluations,
 * [3] environment steps completed. */
SDFC_API sdfc_status sdfc_sim_get_stats(sdfc_sim* sim, int64_t stats[4]);

#ifdef __cplusplus
}
#endif
#endif /* SDFCONTACT_B200_H */
