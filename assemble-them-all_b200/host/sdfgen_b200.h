// sdfgen_b200::make_level_set3 — same signature as sdfgen::make_level_set3
// (externals/sdfgen/makelevelset3.h:15-17), evaluated on the GPU by sdfc_make_level_set3.
// The three call sites of the reference (Body/BodySDFObj.cpp:74, Body/BodyBVHObj.cpp:127,
// Mesh.cpp:90) switch by replacing the namespace; the grid is bit-identical.
#pragma once
#include "makelevelset3.h"

namespace sdfgen_b200 {
void make_level_set3(const std::vector<sdfgen::Vec3ui>& tri, const std::vector<sdfgen::Vec3f>& x, const sdfgen::Vec3f& origin,
                     float dx, float dy, float dz, int nx, int ny, int nz, sdfgen::Array3f& phi, const int exact_band = 1);
}
