#include "sdfgen_b200.h"

#include "B200DeviceCache.h"

namespace sdfgen_b200 {

void make_level_set3(const std::vector<sdfgen::Vec3ui>& tri, const std::vector<sdfgen::Vec3f>& x, const sdfgen::Vec3f& origin,
                     float dx, float dy, float dz, int nx, int ny, int nz, sdfgen::Array3f& phi, const int exact_band) {
    std::vector<uint32_t> t(3 * tri.size());
    std::vector<float> v(3 * x.size());
    for (size_t i = 0; i < tri.size(); ++i) for (int a = 0; a < 3; ++a) t[3 * i + a] = tri[i][a];
    for (size_t i = 0; i < x.size(); ++i) for (int a = 0; a < 3; ++a) v[3 * i + a] = x[i][a];
    const float o[3] = {origin[0], origin[1], origin[2]};
    phi.resize(nx, ny, nz);
    redmax::b200::check(sdfc_make_level_set3(redmax::b200::cache().ctx, (int)tri.size(), t.data(), (int)x.size(), v.data(), o, dx,
                                             dy, dz, nx, ny, nz, exact_band, &phi.a[0]),
                        "sdfgen_b200::make_level_set3");
}

}  // namespace sdfgen_b200
