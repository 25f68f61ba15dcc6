// See ForceGeneralSDFContactB200.h. Host glue only: all arithmetic of the contact
// path runs in libsdfcontact_b200.so (CUDA, sm_100a). No CPU fallback: if the
// library cannot create a device context the constructor's first use throws in
// the reference's convention (throw_error -> std::string, Common.h:58-61).
#include "ForceGeneralSDFContactB200.h"

#include "Body/Body.h"
#include "Body/BodyCuboid.h"
#include "Body/BodySDFObj.h"
#include "Robot.h"
#include "Simulation.h"

#include "B200DeviceCache.h"
#include "B200RobotBatch.h"

#include <vector>

namespace redmax {

using b200::cache;
using b200::check;
using b200::DeviceCache;

sdfc_points* b200::contact_points_of(Body* body, bool general_sdf) {
    DeviceCache& c = cache();
    std::vector<Vector3> cp = body->get_contact_points();  // Body.h:129
    if (general_sdf) {                                     // CollisionDetection.cpp:128-131
        if (auto* cuboid = dynamic_cast<BodyCuboid*>(body)) cp = cuboid->get_general_contact_points();
    }
    std::vector<double> xyz(3 * cp.size());
    for (size_t i = 0; i < cp.size(); ++i) { xyz[3 * i] = cp[i](0); xyz[3 * i + 1] = cp[i](1); xyz[3 * i + 2] = cp[i](2); }
    const uint64_t n = cp.size();
    const uint64_t key = b200::fnv1a(xyz.data(), xyz.size() * sizeof(double), b200::fnv1a(&n, sizeof(n)));
    auto ip = c.pts.find(key);
    if (ip == c.pts.end()) {
        sdfc_points* p = nullptr;
        check(sdfc_points_create(c.ctx, (int)cp.size(), xyz.data(), /*morton_sort=*/1, &p), "B200 contact path: contact point upload");
        ip = c.pts.emplace(key, p).first;
    }
    return ip->second;
}

sdfc_sdf* b200::sdf_of(const BodySDFObj* body) {
    DeviceCache& c = cache();
    // fields of BodySDFObj (Body/BodySDFObj.h:13-16); grid a[i + ni*(j + nj*k)]
    const auto& G = body->_SDF;
    const double meta[13] = {(double)G.ni, (double)G.nj, (double)G.nk, body->_min_box(0), body->_min_box(1), body->_min_box(2),
                             body->_max_box(0), body->_max_box(1), body->_max_box(2), body->_dx, body->_dy, body->_dz,
                             body->_col_th};
    const uint64_t key = b200::fnv1a(G.a.data(), G.a.size() * sizeof(double), b200::fnv1a(meta, sizeof(meta)));
    auto is = c.sdf.find(key);
    if (is == c.sdf.end()) {
        sdfc_sdf* s = nullptr;
        check(sdfc_sdf_create(c.ctx, G.ni, G.nj, G.nk, body->_min_box.data(), body->_max_box.data(), body->_dx, body->_dy,
                              body->_dz, body->_col_th, G.a.data(), &s),
              "ForceGeneralSDFContactB200: SDF upload");
        is = c.sdf.emplace(key, s).first;
    }
    return is->second;
}

ForceGeneralSDFContactB200::ForceGeneralSDFContactB200(Simulation* sim, Body* contact_body, Body* SDF_body, dtype kn,
                                                       dtype kt, dtype mu, dtype damping)
    : ForceGeneralSDFContact(sim, contact_body, SDF_body, kn, kt, mu, damping) {}

void ForceGeneralSDFContactB200::init() { ensure_registered(); }

void ForceGeneralSDFContactB200::ensure_registered() {
    if (_rb) return;
    cache();   // fails loudly without a CUDA device: there is no CPU path behind this force
    _rb = b200::RobotBatch::of(_sim);
    _slot = _rb->add(this);
    _sync_col_th = _SDF_body->_col_th;
}

void ForceGeneralSDFContactB200::record_history(int n_active) {
    if (n_active > 0) {  // CollisionDetection.cpp:143-146
        _contact_body->add_contact_body(_SDF_body->_name);
        _SDF_body->add_contact_body(_contact_body->_name);
    }
}

void ForceGeneralSDFContactB200::computeForce(VectorX& fm, VectorX& fr, bool verbose) {
    auto t0 = clock();
    ensure_registered();
    // device grids are shared between bodies with equal content and col_th is part of their key: a changed
    // threshold makes the batch look its grids up again
    if (_sync_col_th != _SDF_body->_col_th) { _rb->invalidate(); _sync_col_th = _SDF_body->_col_th; }
    _rb->evaluate(false);   // one device call per Robot evaluation, shared by all B200 forces
    record_history(_rb->n_active(_slot));
    sdfc_scatter_force(_contact_body->_index[0], _SDF_body->_index[0], _rb->fm_sub(_slot), fm.data());  // .cpp:60-61
    ++_n_evaluations;
    _time.add("collision detection", 0);
    _time.add("compute force", clock() - t0);
}

void ForceGeneralSDFContactB200::computeForceWithDerivative(VectorX& fm, VectorX& fr, MatrixX& Km, MatrixX& Dm,
                                                            MatrixX& Kr, MatrixX& Dr, bool verbose) {
    auto t0 = clock();
    ensure_registered();
    if (_sync_col_th != _SDF_body->_col_th) { _rb->invalidate(); _sync_col_th = _SDF_body->_col_th; }
    _rb->evaluate(true);
    record_history(_rb->n_active(_slot));
    const int ia = _contact_body->_index[0], ib = _SDF_body->_index[0];
    sdfc_scatter_force(ia, ib, _rb->fm_sub(_slot), fm.data());                                                  // .cpp:88-89
    sdfc_scatter_deriv((int)Km.rows(), ia, ib, _rb->Km_sub(_slot), _rb->Dm_sub(_slot), Km.data(), Dm.data());   // .cpp:91-99 (column-major)
    ++_n_evaluations;
    _time.add("collision detection", 0);
    _time.add("compute force", clock() - t0);
}

void ForceGeneralSDFContactB200::computeForceWithDerivative(VectorX&, VectorX&, MatrixX&, MatrixX&, MatrixX&, MatrixX&,
                                                            MatrixX&, MatrixX&, bool) {
    throw_error("ForceGeneralSDFContactB200: design-parameter derivatives are not implemented on the B200 path");
}

int ForceGeneralSDFContactB200::replace_all(Simulation* sim) {
    int n = 0;
    for (auto& f : sim->_robot->_forces) {
        auto* ref = dynamic_cast<ForceGeneralSDFContact*>(f);
        if (ref == nullptr || dynamic_cast<ForceGeneralSDFContactB200*>(f) != nullptr) continue;
        auto* mine = new ForceGeneralSDFContactB200(sim, ref->_contact_body, ref->_SDF_body, ref->_kn, ref->_kt, ref->_mu,
                                                    ref->_damping);
        mine->_scale = ref->_scale;
        mine->init();
        f = mine;  // forces are leaked by the reference as well (Robot.h:20)
        ++n;
    }
    return n;
}

}  // namespace redmax
