// ForceGeneralSDFContactB200 — the reference-side binding of the B200 contact path.
//
// A redmax::Force plug-in with the constructor, setters and accumulate semantics of
// the reference's ForceGeneralSDFContact (simulation/redmax/Force/
// ForceGeneralSDFContact.{h,cpp}), whose two evaluation entry points
//     computeForce(fm, fr)                               (.cpp:46-67)
//     computeForceWithDerivative(fm, fr, Km, Dm, Kr, Dr) (.cpp:69-105)
// run on the GPU through the C-ABI of include/sdfcontact_b200.h instead of the
// CPU loops (collision_detection_general_SDF + the protected per-contact kernels).
//
// It DERIVES from the reference class, so everything that addresses the force
// through the reference type keeps working unchanged: Robot::set_contact_scale's
// dynamic_cast (Robot.cpp:462-472), the public _kn/_kt/_mu/_damping/_scale members
// and their (non-virtual) setters. This file compiles against the reference's own
// headers; it is built by oracle/ref_build/Makefile (target b200) and is what a
// maintainer adds to simulation/redmax/Force/ (see INTEGRATION.md).
#pragma once
#include "Force/ForceGeneralSDFContact.h"

#include <memory>

namespace redmax {

namespace b200 { class RobotBatch; }

class ForceGeneralSDFContactB200 : public ForceGeneralSDFContact {
public:
    ForceGeneralSDFContactB200(Simulation* sim, Body* contact_body, Body* SDF_body,
                               dtype kn = 1., dtype kt = 0., dtype mu = 0., dtype damping = 0.);

    // registers with the Simulation's RobotBatch (device uploads happen at the first evaluation, once per
    // distinct grid / point set); Robot::init calls it after the first Robot::update (Robot.cpp:156-159)
    void init();

    void computeForce(VectorX& fm, VectorX& fr, bool verbose = false);
    void computeForceWithDerivative(VectorX& fm, VectorX& fr, MatrixX& Km, MatrixX& Dm, MatrixX& Kr, MatrixX& Dr,
                                    bool verbose = false);
    // design-parameter derivatives (.cpp:107-145) are not on the planners' path
    // (they need backward_flag, Simulation.cpp:1262): refuse loudly
    void computeForceWithDerivative(VectorX& fm, VectorX& fr, MatrixX& Km, MatrixX& Dm, MatrixX& Kr, MatrixX& Dr,
                                    MatrixX& dfm_dp, MatrixX& dfr_dp, bool verbose = false);

    // Replace every ForceGeneralSDFContact already registered in sim's Robot by a B200
    // one (same bodies and parameters) and init() it. Returns how many were replaced.
    // For hosts that cannot touch Simulation_Constructor.cpp.
    static int replace_all(Simulation* sim);

    long long _n_evaluations = 0;  // GPU evaluations issued by this force (diagnostics)

private:
    void ensure_registered();
    void record_history(int n_active);

    // every B200 force of this Simulation is evaluated in one device call (B200RobotBatch.h)
    std::shared_ptr<b200::RobotBatch> _rb;
    int _slot = -1;
    dtype _sync_col_th = 0;
};

}  // namespace redmax
