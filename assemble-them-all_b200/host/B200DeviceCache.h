// B200DeviceCache — one device context per process and the geometry already uploaded to it,
// shared by every B200 Force plug-in (ForceGeneralSDFContactB200, ForceGroundContactB200)
// and by the sdfgen shim. Host glue compiled against the reference's headers.
#pragma once
#include "Common.h"
#include "sdfcontact_b200.h"

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

struct DeviceCache {
    sdfc_context* ctx = nullptr;
    std::map<const BodySDFObj*, sdfc_sdf*> sdf;
    std::map<const Body*, sdfc_points*> pts;
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

// Body::_contact_points (Body.h:36,129) uploaded once per body
sdfc_points* contact_points_of(Body* body);

}  // namespace b200
}  // namespace redmax
