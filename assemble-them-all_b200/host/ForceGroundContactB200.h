// ForceGroundContactB200 — reference-side binding of the CUDA ground contact (sdfc_ground_*).
//
// Derives from the reference's ForceGroundContact (simulation/redmax/Force/
// ForceGroundContact.{h,cpp}) and overrides its two evaluation entry points
//     computeForce(fm, fr)                               (.cpp:39-57)
//     computeForceWithDerivative(fm, fr, Km, Dm, Kr, Dr) (.cpp:59-81)
// keeping the public _E_g/_kn/_kt/_mu/_damping members, their setters and the += semantics
// on the body's 6x6 diagonal blocks. Mesh bodies only (the planners' and test_joint_sim's
// case); the analytic sphere branch of collision_detection_ground_object is not on this path.
#pragma once
#include "Force/ForceGroundContact.h"

#include <memory>

struct sdfc_ground;
struct sdfc_points;

namespace redmax {

namespace b200 { class RobotBatch; }

class ForceGroundContactB200 : public ForceGroundContact {
public:
    ForceGroundContactB200(Simulation* sim, Body* contact_body, Matrix4 E_g, dtype kn = 1., dtype kt = 0., dtype mu = 0.,
                           dtype damping = 0.);
    void init();
    void computeForce(VectorX& fm, VectorX& fr, bool verbose = false);
    void computeForceWithDerivative(VectorX& fm, VectorX& fr, MatrixX& Km, MatrixX& Dm, MatrixX& Kr, MatrixX& Dr,
                                    bool verbose = false);
    void computeForceWithDerivative(VectorX& fm, VectorX& fr, MatrixX& Km, MatrixX& Dm, MatrixX& Kr, MatrixX& Dr,
                                    MatrixX& dfm_dp, MatrixX& dfr_dp, bool verbose = false);
    long long _n_evaluations = 0;

private:
    void ensure_uploaded();
    void sync_parameters();
    bool evaluate_batched(bool deriv, const double** out, int* n_active);
    // with SDF contact forces in the same Simulation the ground contacts ride in their device call (B200RobotBatch.h)
    std::shared_ptr<b200::RobotBatch> _rb;
    int _slot = -1;
    sdfc_ground* _ground = nullptr;
    Matrix4 _sync_E_g;
    dtype _sync_kn = 0, _sync_kt = 0, _sync_mu = 0, _sync_damping = 0;
    bool _synced = false;
};

}  // namespace redmax
