// See B200RobotBatch.h. Host glue compiled against the reference's headers.
#include "B200RobotBatch.h"

#include "Body/Body.h"
#include "Body/BodySDFObj.h"
#include "ForceGeneralSDFContactB200.h"
#include "ForceGroundContactB200.h"
#include "Simulation.h"

#include <cstring>
#include <sstream>

namespace redmax {
namespace b200 {

namespace {
std::map<Simulation*, std::weak_ptr<RobotBatch>>& assembling() {
    static std::map<Simulation*, std::weak_ptr<RobotBatch>> m;
    return m;
}
std::map<std::string, PoolEntry>& pool() {
    static std::map<std::string, PoolEntry> m;
    return m;
}
}  // namespace

std::shared_ptr<RobotBatch> RobotBatch::of(Simulation* sim) {
    auto& m = assembling();
    auto it = m.find(sim);
    if (it != m.end()) {
        std::shared_ptr<RobotBatch> rb = it->second.lock();
        // a sealed batch at this address belongs to an earlier Simulation (the address was reused) or has
        // already evaluated: forces that register now start a new one
        if (rb && !rb->_sealed) return rb;
    }
    std::shared_ptr<RobotBatch> rb(new RobotBatch());
    m[sim] = rb;
    return rb;
}

int RobotBatch::body_slot(Body* b) {
    for (size_t i = 0; i < _bodies.size(); ++i)
        if (_bodies[i] == b) return (int)i;
    _bodies.push_back(b);
    return (int)_bodies.size() - 1;
}

int RobotBatch::add(ForceGeneralSDFContactB200* f) {
    if (_sealed) throw_error("B200RobotBatch: a force was added after the first evaluation");
    _sdf.push_back(f);
    _pa.push_back(body_slot(f->_contact_body));
    _pb.push_back(body_slot(f->_SDF_body));
    return (int)_sdf.size() - 1;
}

int RobotBatch::add(ForceGroundContactB200* f) {
    if (_sealed) throw_error("B200RobotBatch: a force was added after the first evaluation");
    _ground.push_back(f);
    _gb.push_back(body_slot(f->_contact_body));
    return (int)_ground.size() - 1;
}

void RobotBatch::seal() {
    _sealed = true;
    DeviceCache& c = cache();
    // structure key: which device point sets / grids sit in which body slots
    std::ostringstream key;
    key << _bodies.size() << "|";
    std::vector<sdfc_points*> pts(_sdf.size());
    std::vector<sdfc_sdf*> sdf(_sdf.size());
    for (size_t i = 0; i < _sdf.size(); ++i) {
        pts[i] = contact_points_of(_sdf[i]->_contact_body, /*general_sdf=*/true);
        sdf[i] = sdf_of(_sdf[i]->_SDF_body);
        key << pts[i] << "," << sdf[i] << "," << _pa[i] << "," << _pb[i] << ";";
    }
    key << "|";
    std::vector<sdfc_points*> gpts(_ground.size());
    for (size_t i = 0; i < _ground.size(); ++i) {
        gpts[i] = contact_points_of(_ground[i]->_contact_body);
        key << gpts[i] << "," << _gb[i] << ";";
    }
    PoolEntry& e = pool()[key.str()];
    if (e.batch == nullptr) {
        double eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        for (size_t i = 0; i < _sdf.size(); ++i) {
            sdfc_pair* p = nullptr;
            check(sdfc_pair_create(c.ctx, pts[i], sdf[i], 1.0, 0.0, 0.0, 0.0, &p), "B200RobotBatch: pair");
            e.pairs.push_back(p);
        }
        for (size_t i = 0; i < _ground.size(); ++i) {
            sdfc_ground* g = nullptr;
            check(sdfc_ground_create(c.ctx, gpts[i], eye, 1.0, 0.0, 0.0, 0.0, &g), "B200RobotBatch: ground");
            e.grounds.push_back(g);
        }
        check(sdfc_batch_create(c.ctx, 1, (int)_bodies.size(), (int)_sdf.size(), e.pairs.data(), _pa.data(), _pb.data(), &e.batch),
              "B200RobotBatch: batch");
        if (!_ground.empty())
            check(sdfc_batch_set_grounds(e.batch, (int)_ground.size(), e.grounds.data(), _gb.data()), "B200RobotBatch: grounds");
    }
    _entry = &e;
    const size_t np = _sdf.size(), ng = _ground.size(), nb = _bodies.size();
    _E.assign(16 * nb, 0.0); _phi.assign(6 * nb, 0.0); _key.assign(22 * nb, 0.0);
    _fm.assign(12 * np, 0.0); _Km.assign(144 * np, 0.0); _Dm.assign(144 * np, 0.0); _na.assign(np, 0);
    _g.assign(78 * ng, 0.0); _gna.assign(ng, 0);
    _key_deriv = -1;
}

void RobotBatch::evaluate(bool deriv) {
    if (_entry == nullptr) { _sealed = false; seal(); }
    const size_t nb = _bodies.size();
    for (size_t b = 0; b < nb; ++b) {
        const Matrix4 E = _bodies[b]->_E_0i;    // column-major, as the C-ABI expects
        const Vector6 phi = _bodies[b]->_phi;
        std::memcpy(_E.data() + 16 * b, E.data(), 16 * sizeof(double));
        std::memcpy(_phi.data() + 6 * b, phi.data(), 6 * sizeof(double));
    }
    // the members' parameters are part of the evaluation's identity (public members, changed without notice)
    _params.clear();
    for (ForceGeneralSDFContactB200* f : _sdf) {
        const double v[7] = {f->_kn, f->_kt, f->_mu, f->_damping, f->_scale, f->_contact_body->_contact_scale,
                             f->_SDF_body->_col_th};
        _params.insert(_params.end(), v, v + 7);
    }
    for (ForceGroundContactB200* f : _ground) {
        _params.insert(_params.end(), f->_E_g.data(), f->_E_g.data() + 16);
        const double v[4] = {f->_kn, f->_kt, f->_mu, f->_damping};
        _params.insert(_params.end(), v, v + 4);
    }
    if (_key_deriv == (deriv ? 1 : 0) && std::memcmp(_key.data(), _E.data(), 16 * nb * sizeof(double)) == 0 &&
        std::memcmp(_key.data() + 16 * nb, _phi.data(), 6 * nb * sizeof(double)) == 0 && _params == _key_params)
        return;   // another force of the same Robot evaluation
    // parameters of every member (public members with non-virtual setters in the reference: read at evaluation time)
    for (size_t i = 0; i < _sdf.size(); ++i) {
        ForceGeneralSDFContactB200* f = _sdf[i];
        sdfc_pair* p = _entry->pairs[i];
        check(sdfc_pair_set_stiffness(p, f->_kn, f->_kt), "set_stiffness");
        check(sdfc_pair_set_friction(p, f->_mu), "set_friction");
        check(sdfc_pair_set_damping(p, f->_damping), "set_damping");
        check(sdfc_pair_set_scale(p, f->_scale), "set_scale");
        check(sdfc_pair_set_contact_scale(p, f->_contact_body->_contact_scale), "set_contact_scale");
    }
    for (size_t i = 0; i < _ground.size(); ++i) {
        ForceGroundContactB200* f = _ground[i];
        sdfc_ground* g = _entry->grounds[i];
        check(sdfc_ground_set_transform(g, f->_E_g.data()), "set_transform");
        check(sdfc_ground_set_stiffness(g, f->_kn, f->_kt), "set_stiffness");
        check(sdfc_ground_set_friction(g, f->_mu), "set_friction");
        check(sdfc_ground_set_damping(g, f->_damping), "set_damping");
    }
    check(sdfc_batch_eval_blocks(_entry->batch, _E.data(), _phi.data(), deriv ? 1 : 0, _fm.data(), _Km.data(), _Dm.data(), _na.data(),
                                 _ground.empty() ? nullptr : _g.data(), _ground.empty() ? nullptr : _gna.data()),
          "B200RobotBatch: evaluation");
    ++device_evaluations;
    std::memcpy(_key.data(), _E.data(), 16 * nb * sizeof(double));
    std::memcpy(_key.data() + 16 * nb, _phi.data(), 6 * nb * sizeof(double));
    _key_deriv = deriv ? 1 : 0;
    _key_params = _params;
}

}  // namespace b200
}  // namespace redmax
