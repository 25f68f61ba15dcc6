// SDFMesh::distance on the GPU — the body of SDFMesh::distance (Mesh.cpp:191-198) replaced by a
// call to sdfc_sdf_distance (batched trilinear sampling, FLT_MAX outside the box; bit-identical to
// SDFMesh::query_SDF, Mesh.cpp:143-189). Mesh::min_distance (Mesh.cpp:17-19) goes through it.
#pragma once
#include "Mesh.h"

namespace redmax {
namespace b200 {
VectorX sdfmesh_distance(SDFMesh& mesh, const Matrix3X& X);
}
}  // namespace redmax
