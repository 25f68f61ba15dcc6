/* sdfgen_oracle.c — CPU restatement of the reference's SDF construction.
 *
 * TEST INFRASTRUCTURE ONLY: the checker for sdfc_make_level_set3 (CUDA). Nothing in the
 * product links or calls this file.
 *
 * Follows /root/reference/simulation/externals/sdfgen/makelevelset3.cpp line by line in
 * plain C (fp32 distances, fp64 grid coordinates and parity counting, the same evaluation
 * order; compile with -ffp-contract=off, the reference's x86-64 build has no FMA):
 *   point_segment_closest_point   makelevelset3.cpp:5-18
 *   point_triangle_distance       makelevelset3.cpp:26-48
 *   check_neighbour / sweep       makelevelset3.cpp:86-124
 *   orientation / point_in_triangle_2d   makelevelset3.cpp:126-162
 *   make_level_set3               makelevelset3.cpp:164-226
 * Vector helpers (dot, mag2, dist): externals/sdfgen/vec.h:214-235,343-348.
 * Grids are x-fastest: a[i + ni*(j + nj*k)] (externals/sdfgen/array3.h:70-80).
 * Pinned against the reference's own make_level_set3 (oracle/_ref/ref_probe) by
 * tests/test_sdfgen.py and tests/golden/sdfgen_small.npz.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float v[3]; } V3;

static V3 sub3(V3 a, V3 b) { V3 r = {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; return r; }
static float dot3(V3 a, V3 b) { float d = a.v[0] * b.v[0]; d += a.v[1] * b.v[1]; d += a.v[2] * b.v[2]; return d; }
static float mag2_3(V3 a) { float l = a.v[0] * a.v[0]; l += a.v[1] * a.v[1]; l += a.v[2] * a.v[2]; return l; }
static float dist3(V3 a, V3 b) {
    float t = a.v[0] - b.v[0];
    float d = t * t;
    t = a.v[1] - b.v[1]; d += t * t;
    t = a.v[2] - b.v[2]; d += t * t;
    return sqrtf(d);
}
/* s*x1 + (1-s)*x2, component-wise in fp32 */
static V3 mix3(float s, V3 x1, V3 x2) {
    const float o = 1 - s;
    V3 r = {{s * x1.v[0] + o * x2.v[0], s * x1.v[1] + o * x2.v[1], s * x1.v[2] + o * x2.v[2]}};
    return r;
}

static V3 seg_closest(V3 x0, V3 x1, V3 x2) {
    V3 dx = sub3(x2, x1);
    double m2 = mag2_3(dx);
    float s12 = (float)(dot3(sub3(x2, x0), dx) / m2);
    if (s12 < 0) s12 = 0;
    else if (s12 > 1) s12 = 1;
    return mix3(s12, x1, x2);
}
static float seg_dist(V3 x0, V3 x1, V3 x2) { return dist3(x0, seg_closest(x0, x1, x2)); }

static float tri_dist(V3 x0, V3 x1, V3 x2, V3 x3) {
    V3 x13 = sub3(x1, x3), x23 = sub3(x2, x3), x03 = sub3(x0, x3);
    float m13 = mag2_3(x13), m23 = mag2_3(x23), d = dot3(x13, x23);
    float det = m13 * m23 - d * d;
    float invdet = 1.f / (det > 1e-30f ? det : 1e-30f);   /* max(a, b): (a < b) ? b : a */
    float a = dot3(x13, x03), b = dot3(x23, x03);
    float w23 = invdet * (m23 * a - d * b);
    float w31 = invdet * (m13 * b - d * a);
    float w12 = 1 - w23 - w31;
    if (w23 >= 0 && w31 >= 0 && w12 >= 0) {
        V3 c = {{w23 * x1.v[0] + w31 * x2.v[0] + w12 * x3.v[0], w23 * x1.v[1] + w31 * x2.v[1] + w12 * x3.v[1],
                 w23 * x1.v[2] + w31 * x2.v[2] + w12 * x3.v[2]}};
        return dist3(x0, c);
    }
    float p, q;
    if (w23 > 0) { p = seg_dist(x0, x1, x2); q = seg_dist(x0, x1, x3); }
    else if (w31 > 0) { p = seg_dist(x0, x1, x2); q = seg_dist(x0, x2, x3); }
    else { p = seg_dist(x0, x1, x3); q = seg_dist(x0, x2, x3); }
    return (q < p) ? q : p;   /* util.h min(a,b): std::min */
}

static int orientation(double x1, double y1, double x2, double y2, double* area) {
    *area = y1 * x2 - x1 * y2;
    if (*area > 0) return 1;
    else if (*area < 0) return -1;
    else if (y2 > y1) return 1;
    else if (y2 < y1) return -1;
    else if (x1 > x2) return 1;
    else if (x1 < x2) return -1;
    else return 0;
}

static int in_tri_2d(double x0, double y0, double x1, double y1, double x2, double y2, double x3, double y3, double* a,
                     double* b, double* c) {
    x1 -= x0; x2 -= x0; x3 -= x0;
    y1 -= y0; y2 -= y0; y3 -= y0;
    int sa = orientation(x2, y2, x3, y3, a);
    if (sa == 0) return 0;
    int sb = orientation(x3, y3, x1, y1, b);
    if (sb != sa) return 0;
    int sc = orientation(x1, y1, x2, y2, c);
    if (sc != sa) return 0;
    double sum = *a + *b + *c;
    *a /= sum; *b /= sum; *c /= sum;
    return 1;
}

static double min3d(double a, double b, double c) { double m = a < b ? a : b; return m < c ? m : c; }
static double max3d(double a, double b, double c) { double m = a > b ? a : b; return m > c ? m : c; }
static int clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }

#define IDX(i, j, k) ((size_t)(i) + (size_t)ni * ((size_t)(j) + (size_t)nj * (size_t)(k)))

static void check_neighbour(const unsigned* tri, const V3* x, float* phi, int* closest, int ni, int nj, V3 gx, int i0,
                            int j0, int k0, int i1, int j1, int k1) {
    const int t = closest[IDX(i1, j1, k1)];
    if (t >= 0) {
        float d = tri_dist(gx, x[tri[3 * t]], x[tri[3 * t + 1]], x[tri[3 * t + 2]]);
        if (d < phi[IDX(i0, j0, k0)]) {
            phi[IDX(i0, j0, k0)] = d;
            closest[IDX(i0, j0, k0)] = t;
        }
    }
}

static void sweep(const unsigned* tri, const V3* x, float* phi, int* closest, const float* origin, float dx, float dy,
                  float dz, int ni, int nj, int nk, int di, int dj, int dk) {
    int i0, i1, j0, j1, k0, k1;
    if (di > 0) { i0 = 1; i1 = ni; } else { i0 = ni - 2; i1 = -1; }
    if (dj > 0) { j0 = 1; j1 = nj; } else { j0 = nj - 2; j1 = -1; }
    if (dk > 0) { k0 = 1; k1 = nk; } else { k0 = nk - 2; k1 = -1; }
    for (int k = k0; k != k1; k += dk) for (int j = j0; j != j1; j += dj) for (int i = i0; i != i1; i += di) {
        V3 gx = {{i * dx + origin[0], j * dy + origin[1], k * dz + origin[2]}};
        check_neighbour(tri, x, phi, closest, ni, nj, gx, i, j, k, i - di, j, k);
        check_neighbour(tri, x, phi, closest, ni, nj, gx, i, j, k, i, j - dj, k);
        check_neighbour(tri, x, phi, closest, ni, nj, gx, i, j, k, i - di, j - dj, k);
        check_neighbour(tri, x, phi, closest, ni, nj, gx, i, j, k, i, j, k - dk);
        check_neighbour(tri, x, phi, closest, ni, nj, gx, i, j, k, i - di, j, k - dk);
        check_neighbour(tri, x, phi, closest, ni, nj, gx, i, j, k, i, j - dj, k - dk);
        check_neighbour(tri, x, phi, closest, ni, nj, gx, i, j, k, i - di, j - dj, k - dk);
    }
}

/* make_level_set3. tri: [n_tri][3], xv: [n_vert][3] fp32, phi_out: ni*nj*nk fp32.
 * closest_out (optional): closest-triangle index per cell after the sweeps. Returns 0. */
int sdfgen_make_level_set3(int n_tri, const unsigned* tri, int n_vert, const float* xv, const float* origin, float dx,
                           float dy, float dz, int ni, int nj, int nk, int exact_band, float* phi, int* closest_out) {
    (void)n_vert;
    const V3* x = (const V3*)xv;
    const size_t n = (size_t)ni * nj * nk;
    const float init = (ni + nj + nk) * (dx + dy + dz);
    for (size_t c = 0; c < n; ++c) phi[c] = init;
    int* closest = (int*)malloc(n * sizeof(int));
    int* count = (int*)calloc(n, sizeof(int));
    if (!closest || !count) { free(closest); free(count); return -1; }
    for (size_t c = 0; c < n; ++c) closest[c] = -1;
    for (int t = 0; t < n_tri; ++t) {
        const unsigned p = tri[3 * t], q = tri[3 * t + 1], r = tri[3 * t + 2];
        double fip = ((double)x[p].v[0] - origin[0]) / dx, fjp = ((double)x[p].v[1] - origin[1]) / dy, fkp = ((double)x[p].v[2] - origin[2]) / dz;
        double fiq = ((double)x[q].v[0] - origin[0]) / dx, fjq = ((double)x[q].v[1] - origin[1]) / dy, fkq = ((double)x[q].v[2] - origin[2]) / dz;
        double fir = ((double)x[r].v[0] - origin[0]) / dx, fjr = ((double)x[r].v[1] - origin[1]) / dy, fkr = ((double)x[r].v[2] - origin[2]) / dz;
        int i0 = clampi((int)min3d(fip, fiq, fir) - exact_band, 0, ni - 1), i1 = clampi((int)max3d(fip, fiq, fir) + exact_band + 1, 0, ni - 1);
        int j0 = clampi((int)min3d(fjp, fjq, fjr) - exact_band, 0, nj - 1), j1 = clampi((int)max3d(fjp, fjq, fjr) + exact_band + 1, 0, nj - 1);
        int k0 = clampi((int)min3d(fkp, fkq, fkr) - exact_band, 0, nk - 1), k1 = clampi((int)max3d(fkp, fkq, fkr) + exact_band + 1, 0, nk - 1);
        for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) {
            V3 gx = {{i * dx + origin[0], j * dy + origin[1], k * dz + origin[2]}};
            float d = tri_dist(gx, x[p], x[q], x[r]);
            if (d < phi[IDX(i, j, k)]) { phi[IDX(i, j, k)] = d; closest[IDX(i, j, k)] = t; }
        }
        j0 = clampi((int)ceil(min3d(fjp, fjq, fjr)), 0, nj - 1);
        j1 = clampi((int)floor(max3d(fjp, fjq, fjr)), 0, nj - 1);
        k0 = clampi((int)ceil(min3d(fkp, fkq, fkr)), 0, nk - 1);
        k1 = clampi((int)floor(max3d(fkp, fkq, fkr)), 0, nk - 1);
        for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) {
            double a, b, c;
            if (in_tri_2d(j, k, fjp, fkp, fjq, fkq, fjr, fkr, &a, &b, &c)) {
                double fi = a * fip + b * fiq + c * fir;
                int ii = (int)ceil(fi);
                if (ii < 0) ++count[IDX(0, j, k)];
                else if (ii < ni) ++count[IDX(ii, j, k)];
            }
        }
    }
    for (int pass = 0; pass < 2; ++pass) {
        sweep(tri, x, phi, closest, origin, dx, dy, dz, ni, nj, nk, +1, +1, +1);
        sweep(tri, x, phi, closest, origin, dx, dy, dz, ni, nj, nk, -1, -1, -1);
        sweep(tri, x, phi, closest, origin, dx, dy, dz, ni, nj, nk, +1, +1, -1);
        sweep(tri, x, phi, closest, origin, dx, dy, dz, ni, nj, nk, -1, -1, +1);
        sweep(tri, x, phi, closest, origin, dx, dy, dz, ni, nj, nk, +1, -1, +1);
        sweep(tri, x, phi, closest, origin, dx, dy, dz, ni, nj, nk, -1, +1, -1);
        sweep(tri, x, phi, closest, origin, dx, dy, dz, ni, nj, nk, +1, -1, -1);
        sweep(tri, x, phi, closest, origin, dx, dy, dz, ni, nj, nk, -1, +1, +1);
    }
    for (int k = 0; k < nk; ++k) for (int j = 0; j < nj; ++j) {
        int total = 0;
        for (int i = 0; i < ni; ++i) {
            total += count[IDX(i, j, k)];
            if (total % 2 == 1) phi[IDX(i, j, k)] = -phi[IDX(i, j, k)];
        }
    }
    if (closest_out) memcpy(closest_out, closest, n * sizeof(int));
    free(closest);
    free(count);
    return 0;
}

/* Grid layout of BodySDFObj::precompute_SDF (Body/BodySDFObj.cpp:38-70) and SDFMesh
 * (Mesh.cpp:54-100): per-axis spacing min(dx, bbox/res), one cell of padding, fp32 box.
 * V: [n][3] fp64 body-frame vertices. Outputs: spacing[3] (double), min_box / max_box (fp32),
 * dims[3]. */
void sdfgen_grid_layout(int n, const double* V, double dx, int res, double spacing[3], float min_box[3], float max_box[3],
                        int dims[3]) {
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { min_box[a] = 3.402823466e+38F; max_box[a] = -3.402823466e+38F; lo[a] = V[a]; hi[a] = V[a]; }
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            const float p = (float)V[3 * i + a];
            if (p < min_box[a]) min_box[a] = p;
            if (p > max_box[a]) max_box[a] = p;
            if (V[3 * i + a] < lo[a]) lo[a] = V[3 * i + a];
            if (V[3 * i + a] > hi[a]) hi[a] = V[3 * i + a];
        }
    for (int a = 0; a < 3; ++a) {
        const double ext = (hi[a] - lo[a]) / res;      /* _bounding_box is the fp64 vertex box */
        spacing[a] = dx < ext ? dx : ext;
        const float unit = (float)spacing[a];
        min_box[a] -= 1.f * unit;
        max_box[a] += 1.f * unit;
        const float half = (float)(0.5 * spacing[a]);
        const float num = max_box[a] - min_box[a] + half;
        dims[a] = (int)(unsigned)(num / (float)spacing[a]) + 1;
    }
}
