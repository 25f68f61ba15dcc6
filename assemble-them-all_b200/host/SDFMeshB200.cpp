#include "SDFMeshB200.h"

#include "B200DeviceCache.h"

#include <map>

namespace redmax {
namespace b200 {

namespace {
struct Uploaded {
    sdfc_sdf* sdf = nullptr;
    const double* data = nullptr;  // grid storage the upload was made from
    int ni = 0, nj = 0, nk = 0;
    double probe = 0;              // cheap content check: meshes are created and destroyed per planner
};
double grid_probe(const SDFMesh& m) {
    const size_t n = (size_t)m._SDF.ni * m._SDF.nj * m._SDF.nk;
    double s = 0;
    for (size_t i = 0; i < n; i += (n / 61) + 1) s += m._SDF.a[i] * (double)(i % 7 + 1);
    return s + m._dx + 3 * m._min_box[0];
}
}  // namespace

VectorX sdfmesh_distance(SDFMesh& mesh, const Matrix3X& X) {
    static std::map<const SDFMesh*, Uploaded> cache_;
    Uploaded& u = cache_[&mesh];
    const double pr = grid_probe(mesh);
    if (u.sdf == nullptr || u.data != &mesh._SDF.a[0] || u.ni != mesh._SDF.ni || u.nj != mesh._SDF.nj || u.nk != mesh._SDF.nk ||
        u.probe != pr) {
        if (u.sdf) sdfc_sdf_destroy(u.sdf);
        u = Uploaded();
        check(sdfc_sdf_create(cache().ctx, mesh._SDF.ni, mesh._SDF.nj, mesh._SDF.nk, mesh._min_box.data(), mesh._max_box.data(),
                              mesh._dx, mesh._dy, mesh._dz, /*col_th=*/0.0, &mesh._SDF.a[0], &u.sdf),
              "SDFMesh (B200): grid upload");
        u.data = &mesh._SDF.a[0]; u.ni = mesh._SDF.ni; u.nj = mesh._SDF.nj; u.nk = mesh._SDF.nk; u.probe = pr;
    }
    VectorX d(X.cols());
    if (X.cols() > 0)  // Matrix3X is column-major: xyz of a point are contiguous, the AoS the C-ABI takes
        check(sdfc_sdf_distance(u.sdf, (int64_t)X.cols(), X.data(), d.data()), "SDFMesh::distance (B200)");
    return d;
}

}  // namespace b200
}  // namespace redmax
