// B200RobotBatch — every B200 contact force of one Simulation evaluated in ONE device call.
//
// The reference's plug-in interface is per force: Robot::computeForce[WithDerivative] (Robot.cpp:513-551) calls
// each Force in turn, after Robot::update has refreshed every Body::_E_0i / _phi. A force that uploads its two
// poses, launches its kernels, downloads its blocks and synchronises on its own pays that latency once per
// force per Newton / line-search evaluation (38 times per evaluation in a 20-part assembly).
// Here the forces of a Simulation share a RobotBatch: the FIRST force asked in an evaluation gathers the poses
// of all member bodies, pushes every member's parameters, runs one sdfc_batch_eval_blocks (one pinned upload,
// detection + force kernels over all pairs and grounds, one download, one synchronisation) and caches the
// per-force blocks; the other forces of the same evaluation (same poses, same derivative flag) only add their
// own blocks. Accumulation order and arithmetic per force are unchanged.
//
// Device objects are pooled by structure (which point sets / grids / body slots), not by Simulation: the planners
// build a new Simulation per candidate and the reference never frees its forces, so per-Simulation device
// buffers would accumulate without bound.
#pragma once
#include "B200DeviceCache.h"

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace redmax {

class Simulation;
class ForceGeneralSDFContactB200;
class ForceGroundContactB200;

namespace b200 {

struct PoolEntry {
    std::vector<sdfc_pair*> pairs;
    std::vector<sdfc_ground*> grounds;
    sdfc_batch* batch = nullptr;
};

class RobotBatch {
public:
    // Called from Force::init(): the batch this Simulation's forces are being collected into.
    static std::shared_ptr<RobotBatch> of(Simulation* sim);
    int add(ForceGeneralSDFContactB200* f);   // -> slot among the SDF contact forces
    int add(ForceGroundContactB200* f);       // -> slot among the ground contact forces
    void invalidate() { _entry = nullptr; _key_deriv = -1; }   // a member's geometry / threshold changed
    bool batched() const { return !_sdf.empty(); }             // ground contacts alone keep their per-force path

    // Evaluates all member forces at the bodies' current poses, unless the last evaluation had the same
    // poses and derivative flag (the other forces of the same Robot evaluation).
    void evaluate(bool deriv);
    const double* fm_sub(int slot) const { return _fm.data() + 12 * (size_t)slot; }
    const double* Km_sub(int slot) const { return _Km.data() + 144 * (size_t)slot; }
    const double* Dm_sub(int slot) const { return _Dm.data() + 144 * (size_t)slot; }
    int n_active(int slot) const { return _na[slot]; }
    const double* ground_out(int slot) const { return _g.data() + 78 * (size_t)slot; }
    int ground_active(int slot) const { return _gna[slot]; }
    long long device_evaluations = 0;   // diagnostics: sdfc_batch_eval_blocks calls

private:
    void seal();
    int body_slot(Body* b);
    std::vector<ForceGeneralSDFContactB200*> _sdf;
    std::vector<ForceGroundContactB200*> _ground;
    std::vector<Body*> _bodies;
    std::vector<int32_t> _pa, _pb, _gb;
    PoolEntry* _entry = nullptr;
    bool _sealed = false;
    std::vector<double> _E, _phi, _key, _params, _key_params;
    int _key_deriv = -1;
    std::vector<double> _fm, _Km, _Dm, _g;
    std::vector<int32_t> _na, _gna;
};

}  // namespace b200
}  // namespace redmax
