// Host build of the stepper's per-environment arithmetic (assemble-them-all_b200/csrc/sdfc_step_math.cuh).
// TEST INFRASTRUCTURE: lets the CPU suite compare the warp-cooperative restatement with the numpy
// oracle (oracle/sim_oracle.py) and check that its statements are order independent (dir = +1 / -1).
#include "../../assemble-them-all_b200/csrc/sdfc_step_math.cuh"

#include <cstring>
#include <vector>

using namespace sdfc;

extern "C" {

// bodies: per body joint_type, E_pj_0[16], E_ji[16] (column-major), inertia[6], mass.
// x: candidate state; q0, qdot0: previous state. fm_sub/Km_sub/Dm_sub per pair (column-major 12x12).
// Outputs: g[nr], H[nr*nr] row-major (deriv), E[n_bodies*16] column-major, phi[n_bodies*6], dx[nr] = solve(H, -g) (deriv).
int step_eval(int n_bodies, const int* joint_type, const double* E_pj_0, const double* E_ji, const double* inertia,
              const double* mass, double h, const double* grav, const double* x, const double* q0, const double* qdot0,
              const double* fext, int n_pairs, const int* pair_a, const int* pair_b, const double* fm_sub,
              const double* Km_sub, const double* Dm_sub, int deriv, int dir, double* g, double* H, double* E, double* phi,
              double* dx) {
    std::vector<BodyDev> bodies(n_bodies);
    int nr = 0, nd = 0;
    for (int b = 0; b < n_bodies; ++b) {
        make_body_dev(joint_type[b], E_pj_0 + 16 * b, E_ji + 16 * b, inertia + 6 * b, mass[b], nr, nd, bodies[b]);
        nr += bodies[b].ndof;
        if (bodies[b].ndof > 0) ++nd;
    }
    if (nr > kStepCapR || nd > kStepCapDyn) return -1;
    EnvW* W = new EnvW();
    std::memset(W, 0xFF, sizeof(EnvW));   // NaN-poison: anything read before it is written shows up
    WarpExec wx{0, dir};
    for (int j = 0; j < nr; ++j) { W->xc[j] = x[j]; W->qdc[j] = (x[j] - q0[j]) / h; }
    EvalIn in{};
    in.bodies = bodies.data(); in.n_bodies = n_bodies; in.n_pairs = n_pairs; in.n_dyn = nd; in.nr = nr; in.n_grounds = 0;
    in.pair_a = pair_a; in.pair_b = pair_b; in.ground_body = nullptr;
    in.h = h; in.grav[0] = grav[0]; in.grav[1] = grav[1]; in.grav[2] = grav[2];
    in.q0 = q0; in.qdot0 = qdot0; in.fext = fext; in.fm_sub = fm_sub; in.Km_sub = Km_sub; in.Dm_sub = Dm_sub; in.g_out = nullptr;
    w_evaluate(wx, in, deriv != 0, *W);
    for (int i = 0; i < nr; ++i) g[i] = W->g[i];
    if (deriv) {
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < nr; ++j) H[i * nr + j] = W->H[i * kStepCapR + j];
        for (int i = 0; i < nr; ++i) W->rhs[i] = -W->g[i];
        w_lu_solve(wx, W->H, W->rhs, nr, &W->piv);
        for (int i = 0; i < nr; ++i) dx[i] = W->rhs[i];
    }
    w_poses(wx, bodies.data(), n_bodies, *W, E, phi);
    delete W;
    return nr;
}

// JointSphericalExp::reparam on (r, rd) in place.
void step_reparam(double* r, double* rd, int dir) {
    EnvW* W = new EnvW();
    std::memset(W, 0xFF, sizeof(EnvW));
    WarpExec wx{0, dir};
    w_reparam(wx, r, rd, *W);
    delete W;
}

}  // extern "C"
