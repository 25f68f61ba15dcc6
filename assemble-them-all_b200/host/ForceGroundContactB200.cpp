// See ForceGroundContactB200.h. Host glue only; the arithmetic runs in libsdfcontact_b200.so.
#include "ForceGroundContactB200.h"

#include "B200DeviceCache.h"
#include "B200RobotBatch.h"
#include "Body/Body.h"
#include "Body/BodySphere.h"
#include "Simulation.h"

namespace redmax {

using b200::check;

ForceGroundContactB200::ForceGroundContactB200(Simulation* sim, Body* contact_body, Matrix4 E_g, dtype kn, dtype kt, dtype mu,
                                               dtype damping)
    : ForceGroundContact(sim, contact_body, E_g, kn, kt, mu, damping) {}

void ForceGroundContactB200::init() {
    if (dynamic_cast<BodySphere*>(_contact_body) != nullptr)
        throw_error("ForceGroundContactB200: analytic sphere bodies are not supported on the B200 path");
    b200::cache();   // fails loudly without a CUDA device
    if (!_rb) {
        _rb = b200::RobotBatch::of(_sim);
        _slot = _rb->add(this);
    }
}

// One device call per Robot evaluation when the Simulation also has SDF contact forces; -> false when this
// force has to evaluate on its own (no SDF contact force in the Simulation).
bool ForceGroundContactB200::evaluate_batched(bool deriv, const double** out, int* n_active) {
    if (!_rb) init();
    if (!_rb->batched()) return false;
    _rb->evaluate(deriv);
    *out = _rb->ground_out(_slot);
    *n_active = _rb->ground_active(_slot);
    return true;
}

void ForceGroundContactB200::ensure_uploaded() {
    if (_ground != nullptr) return;
    if (dynamic_cast<BodySphere*>(_contact_body) != nullptr)
        throw_error("ForceGroundContactB200: analytic sphere bodies are not supported on the B200 path");
    sdfc_points* pts = b200::contact_points_of(_contact_body);
    check(sdfc_ground_create(b200::cache().ctx, pts, _E_g.data(), _kn, _kt, _mu, _damping, &_ground),
          "ForceGroundContactB200: create");
    _sync_E_g = _E_g; _sync_kn = _kn; _sync_kt = _kt; _sync_mu = _mu; _sync_damping = _damping;
    _synced = true;
}

void ForceGroundContactB200::sync_parameters() {
    // the reference's setters are non-virtual and the members public: detect changes here
    if (_synced && _sync_kn == _kn && _sync_kt == _kt && _sync_mu == _mu && _sync_damping == _damping && _sync_E_g == _E_g) return;
    check(sdfc_ground_set_transform(_ground, _E_g.data()), "set_transform");
    check(sdfc_ground_set_stiffness(_ground, _kn, _kt), "set_stiffness");
    check(sdfc_ground_set_friction(_ground, _mu), "set_friction");
    check(sdfc_ground_set_damping(_ground, _damping), "set_damping");
    _sync_E_g = _E_g; _sync_kn = _kn; _sync_kt = _kt; _sync_mu = _mu; _sync_damping = _damping;
    _synced = true;
}

void ForceGroundContactB200::computeForce(VectorX& fm, VectorX& fr, bool verbose) {
    auto t0 = clock();
    double f[6];
    int32_t n_active = 0;
    const double* blk = nullptr;
    if (evaluate_batched(false, &blk, &n_active)) {
        for (int r = 0; r < 6; ++r) f[r] = blk[r];
    } else {
        ensure_uploaded();
        sync_parameters();
        const Matrix4 E = _contact_body->_E_0i;
        const Vector6 phi = _contact_body->_phi;
        check(sdfc_ground_force(_ground, E.data(), phi.data(), f, &n_active), "ForceGroundContactB200::computeForce");
    }
    if (n_active > 0) _contact_body->add_contact_body("ground");  // CollisionDetection.cpp:67
    const int i0 = _contact_body->_index[0];
    for (int r = 0; r < 6; ++r) fm(i0 + r) += f[r];               // .cpp:129
    ++_n_evaluations;
    _time.add("collision detection", 0);
    _time.add("compute force", clock() - t0);
}

void ForceGroundContactB200::computeForceWithDerivative(VectorX& fm, VectorX& fr, MatrixX& Km, MatrixX& Dm, MatrixX& Kr,
                                                        MatrixX& Dr, bool verbose) {
    auto t0 = clock();
    double f[6], K[36], D[36];
    int32_t n_active = 0;
    const double* blk = nullptr;
    if (evaluate_batched(true, &blk, &n_active)) {
        for (int r = 0; r < 6; ++r) f[r] = blk[r];
        for (int i = 0; i < 36; ++i) { K[i] = blk[6 + i]; D[i] = blk[42 + i]; }
    } else {
        ensure_uploaded();
        sync_parameters();
        const Matrix4 E = _contact_body->_E_0i;
        const Vector6 phi = _contact_body->_phi;
        check(sdfc_ground_force_deriv(_ground, E.data(), phi.data(), f, K, D, &n_active),
              "ForceGroundContactB200::computeForceWithDerivative");
    }
    if (n_active > 0) _contact_body->add_contact_body("ground");
    const int i0 = _contact_body->_index[0];
    for (int r = 0; r < 6; ++r) fm(i0 + r) += f[r];
    for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) {  // column-major 6x6 blocks (.cpp:186-189,209-210,237-238)
            Km(i0 + r, i0 + c) += K[r + 6 * c];
            Dm(i0 + r, i0 + c) += D[r + 6 * c];
        }
    ++_n_evaluations;
    _time.add("collision detection", 0);
    _time.add("compute force", clock() - t0);
}

void ForceGroundContactB200::computeForceWithDerivative(VectorX&, VectorX&, MatrixX&, MatrixX&, MatrixX&, MatrixX&, MatrixX&,
                                                        MatrixX&, bool) {
    throw_error("ForceGroundContactB200: design-parameter derivatives are not implemented on the B200 path");
}

}  // namespace redmax
