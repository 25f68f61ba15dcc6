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

void ForceGeneralSDFContactB200::init() { ensure_uploaded(); }

void ForceGeneralSDFContactB200::ensure_uploaded() {
    if (_pair != nullptr) return;
    _ctx = cache().ctx;
    _sdf = b200::sdf_of(_SDF_body);
    _pts = b200::contact_points_of(_contact_body, /*general_sdf=*/true);
    check(sdfc_pair_create(_ctx, _pts, _sdf, _kn, _kt, _mu, _damping, &_pair), "ForceGeneralSDFContactB200: pair");
    _sync_col_th = _SDF_body->_col_th;
    _synced = false;
}

void ForceGeneralSDFContactB200::sync_parameters() {
    const dtype cs = _contact_body->_contact_scale, cth = _SDF_body->_col_th;
    if (_synced && _sync_kn == _kn && _sync_kt == _kt && _sync_mu == _mu && _sync_damping == _damping &&
        _sync_scale == _scale && _sync_contact_scale == cs && _sync_col_th == cth)
        return;
    check(sdfc_pair_set_stiffness(_pair, _kn, _kt), "set_stiffness");
    check(sdfc_pair_set_friction(_pair, _mu), "set_friction");
    check(sdfc_pair_set_damping(_pair, _damping), "set_damping");
    check(sdfc_pair_set_scale(_pair, _scale), "set_scale");
    check(sdfc_pair_set_contact_scale(_pair, cs), "set_contact_scale");
    if (_sync_col_th != cth) {
        // device grids are shared between bodies with equal content (B200DeviceCache.h) and col_th is part
        // of the key: a changed threshold re-points this force at the matching grid instead of mutating a
        // shared one
        check(sdfc_pair_destroy(_pair), "pair_destroy");
        _pair = nullptr;
        _sdf = b200::sdf_of(_SDF_body);
        check(sdfc_pair_create(_ctx, _pts, _sdf, _kn, _kt, _mu, _damping, &_pair), "ForceGeneralSDFContactB200: pair");
    }
    _sync_kn = _kn; _sync_kt = _kt; _sync_mu = _mu; _sync_damping = _damping; _sync_scale = _scale;
    _sync_contact_scale = cs; _sync_col_th = cth;
    _synced = true;
}

void ForceGeneralSDFContactB200::record_history(int n_active) {
    if (n_active > 0) {  // CollisionDetection.cpp:143-146
        _contact_body->add_contact_body(_SDF_body->_name);
        _SDF_body->add_contact_body(_contact_body->_name);
    }
}

void ForceGeneralSDFContactB200::computeForce(VectorX& fm, VectorX& fr, bool verbose) {
    auto t0 = clock();
    ensure_uploaded();
    sync_parameters();
    double fm_sub[12];
    int32_t n_active = 0;
    const Matrix4 E1 = _contact_body->_E_0i, E2 = _SDF_body->_E_0i;  // column-major, as the C-ABI expects
    const Vector6 phi1 = _contact_body->_phi, phi2 = _SDF_body->_phi;
    check(sdfc_pair_force(_pair, E1.data(), phi1.data(), E2.data(), phi2.data(), fm_sub, &n_active),
          "ForceGeneralSDFContactB200::computeForce");
    record_history(n_active);
    sdfc_scatter_force(_contact_body->_index[0], _SDF_body->_index[0], fm_sub, fm.data());  // .cpp:60-61
    ++_n_evaluations;
    _time.add("collision detection", 0);
    _time.add("compute force", clock() - t0);
}

void ForceGeneralSDFContactB200::computeForceWithDerivative(VectorX& fm, VectorX& fr, MatrixX& Km, MatrixX& Dm,
                                                            MatrixX& Kr, MatrixX& Dr, bool verbose) {
    auto t0 = clock();
    ensure_uploaded();
    sync_parameters();
    double fm_sub[12], Km_sub[144], Dm_sub[144];
    int32_t n_active = 0;
    const Matrix4 E1 = _contact_body->_E_0i, E2 = _SDF_body->_E_0i;
    const Vector6 phi1 = _contact_body->_phi, phi2 = _SDF_body->_phi;
    check(sdfc_pair_force_deriv(_pair, E1.data(), phi1.data(), E2.data(), phi2.data(), fm_sub, Km_sub, Dm_sub, &n_active),
          "ForceGeneralSDFContactB200::computeForceWithDerivative");
    record_history(n_active);
    const int ia = _contact_body->_index[0], ib = _SDF_body->_index[0];
    sdfc_scatter_force(ia, ib, fm_sub, fm.data());                                      // .cpp:88-89
    sdfc_scatter_deriv((int)Km.rows(), ia, ib, Km_sub, Dm_sub, Km.data(), Dm.data());   // .cpp:91-99 (column-major)
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
