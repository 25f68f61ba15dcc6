// B200DeviceCache — one device context per process and the geometry already uploaded to it,
// shared by every B200 Force plug-in (ForceGeneralSDFContactB200, ForceGroundContactB200)
// and by the sdfgen shim. Host glue compiled against the reference's headers.
#pragma once
#include "Common.h"
#include "sdfcontact_b200.h"

#include <cstdint>
#include <cstdlib>
#include <map>
#include <string>

namespace redmax {

class Body;
class BodySDFObj;

namespace b200 {

inline void check(sdfc_status st, const char* what) {
    if (st != SDFC_OK) throw_error(std::string(what) + ": " + sdfc_last_error());  // Common.h:58-61 convention
}

// Uploaded geometry is keyed by CONTENT (FNV-1a over the array bytes), not by Body*: the planners
// construct a new Simulation per candidate and the reference never frees its bodies, so an
// address-keyed cache would add one device copy per construction (~40 MB per 100^3 grid) and could
// alias a reused address. Identical parts of successive Simulations share one upload; device
// memory is bounded by the number of DISTINCT grids / point sets of the run.
inline uint64_t fnv1a(const void* data, size_t bytes, uint64_t h = 1469598103934665603ull) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

struct DeviceCache {
    sdfc_context* ctx = nullptr;
    std::map<uint64_t, sdfc_sdf*> sdf;   // key: dims, box, spacing, col_th, grid values
    std::map<uint64_t, sdfc_points*> pts;  // key: point coordinates (+ which point set of the body)
};

// device from SDFC_DEVICE (default 0); created on first use, i.e. after a fork (utils/parallel.py:27)
inline DeviceCache& cache() {
    static DeviceCache c;
    if (c.ctx == nullptr) {
        const char* dev = std::getenv("SDFC_DEVICE");
        check(sdfc_context_create(dev ? std::atoi(dev) : 0, &c.ctx), "B200 contact path: no CUDA device");
    }
    return c;
}

// Body::_contact_points (Body.h:36,129), uploaded once per distinct point set. general_sdf = true
// selects the set collision_detection_general_SDF uses (CollisionDetection.cpp:126-131: a BodyCuboid
// contributes get_general_contact_points(), its 6x6x6 lattice, instead of its corner points).
sdfc_points* contact_points_of(Body* body, bool general_sdf = false);
// BodySDFObj grid fields (Body/BodySDFObj.h:13-16), uploaded once per distinct (grid, col_th)
sdfc_sdf* sdf_of(const BodySDFObj* body);

}  // namespace b200
}  // namespace redmax
