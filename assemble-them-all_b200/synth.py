"""Deterministic synthetic inputs for the SDF-contact hot path.

No datasets are available offline, so every mesh, grid and pose used by the
tests and by bench.py is generated here from fixed seeds:

* watertight peg / perforated-block / multi-part meshes, written as OBJ with
  ``%.9f`` vertices (the reference reads them through tiny_obj_loader,
  simulation/redmax/Body/BodyMeshObj.cpp:46-68);
* the planner's simulation XML (mirrors examples/run_joint_plan.py:140-178)
  and the free-fall XML (mirrors examples/test_joint_sim.py:24-63);
* analytic sphere SDF grids for the C5 query sweep, rounded to fp32 and widened
  like freshly generated grids are (Body/BodySDFObj.cpp:73-83);
* seeded rigid poses / body velocities.

Meshes are produced by a voxel-lattice surface mesher: a solid is a boolean
occupancy over a (possibly non-uniform) axis-aligned lattice, and its boundary
quads are emitted with welded lattice vertices — watertight by construction.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

MESH_SEED = 20221
SWEEP_SEED = 1234


# --------------------------------------------------------------------------
# lattice surface mesher
# --------------------------------------------------------------------------
def _axis(breaks, h):
    """Piecewise-uniform coordinates through every value in `breaks`, spacing <= h."""
    out = [np.array([breaks[0]], dtype=np.float64)]
    for a, b in zip(breaks[:-1], breaks[1:]):
        n = max(1, int(np.ceil((b - a) / h - 1e-9)))
        out.append(a + (b - a) * np.arange(1, n + 1) / n)
    return np.concatenate(out)


def lattice_surface(xs, ys, zs, occ):
    """Boundary surface of the occupied cells of a lattice.

    xs, ys, zs: increasing coordinate arrays of length nx+1, ny+1, nz+1.
    occ: bool array (nx, ny, nz).
    Returns (V[n,3] float64, F[m,3] int32), outward-oriented (CCW from outside).
    """
    occ = np.asarray(occ, dtype=bool)
    nx, ny, nz = occ.shape
    assert len(xs) == nx + 1 and len(ys) == ny + 1 and len(zs) == nz + 1
    pad = np.zeros((nx + 2, ny + 2, nz + 2), dtype=bool)
    pad[1:-1, 1:-1, 1:-1] = occ

    def pid(a, b, c):
        return a + (nx + 1) * (b + (ny + 1) * c)

    quads = []
    # (axis, sign) -> corner offsets in CCW order seen from outside
    corner = {
        (0, +1): [(1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1)],
        (0, -1): [(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0)],
        (1, +1): [(0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)],
        (1, -1): [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1)],
        (2, +1): [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)],
        (2, -1): [(0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 0)],
    }
    for (ax, sg), offs in corner.items():
        sl = [slice(1, -1)] * 3
        sl[ax] = slice(2, None) if sg > 0 else slice(0, -2)
        exposed = occ & ~pad[tuple(sl)]
        i, j, k = np.nonzero(exposed)
        q = np.stack([pid(i + o[0], j + o[1], k + o[2]) for o in offs], axis=1)
        quads.append(q)
    quads = np.concatenate(quads, axis=0)
    used, inv = np.unique(quads.reshape(-1), return_inverse=True)
    quads = inv.reshape(-1, 4)
    a = used % (nx + 1)
    b = (used // (nx + 1)) % (ny + 1)
    c = used // ((nx + 1) * (ny + 1))
    V = np.stack([np.asarray(xs)[a], np.asarray(ys)[b], np.asarray(zs)[c]], axis=1).astype(np.float64)
    F = np.concatenate([quads[:, [0, 1, 2]], quads[:, [0, 2, 3]]], axis=0).astype(np.int32)
    return V, F


def box_mesh(size, center=(0.0, 0.0, 0.0), h=0.1):
    """Axis-aligned box surface with lattice spacing <= h."""
    size = np.asarray(size, dtype=np.float64)
    center = np.asarray(center, dtype=np.float64)
    lo, hi = center - size / 2, center + size / 2
    xs, ys, zs = (_axis([lo[a], hi[a]], h) for a in range(3))
    occ = np.ones((len(xs) - 1, len(ys) - 1, len(zs) - 1), dtype=bool)
    return lattice_surface(xs, ys, zs, occ)


def perforated_block_mesh(size, holes, hole_size, center=(0.0, 0.0, 0.0), h=0.1):
    """Block with square through-holes along z. `holes` = list of (cx, cy)."""
    size = np.asarray(size, dtype=np.float64)
    center = np.asarray(center, dtype=np.float64)
    lo, hi = center - size / 2, center + size / 2
    bx, by = {lo[0], hi[0]}, {lo[1], hi[1]}
    for cx, cy in holes:
        bx |= {cx - hole_size / 2, cx + hole_size / 2}
        by |= {cy - hole_size / 2, cy + hole_size / 2}
    xs, ys = _axis(sorted(bx), h), _axis(sorted(by), h)
    zs = _axis([lo[2], hi[2]], h)
    xc, yc = 0.5 * (xs[:-1] + xs[1:]), 0.5 * (ys[:-1] + ys[1:])
    occ2 = np.ones((len(xc), len(yc)), dtype=bool)
    for cx, cy in holes:
        inx = np.abs(xc - cx) < hole_size / 2
        iny = np.abs(yc - cy) < hole_size / 2
        occ2 &= ~(inx[:, None] & iny[None, :])
    occ = np.repeat(occ2[:, :, None], len(zs) - 1, axis=2)
    return lattice_surface(xs, ys, zs, occ)


def mesh_is_watertight(V, F):
    """Every undirected edge shared by exactly two triangles, with opposite directions."""
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0).astype(np.int64)
    key = e[:, 0] * (len(V) + 1) + e[:, 1]
    rkey = e[:, 1] * (len(V) + 1) + e[:, 0]
    if len(np.unique(key)) != len(key):
        return False
    return bool(np.array_equal(np.sort(key), np.sort(rkey)))


def write_obj(path, V, F):
    with open(path, "w") as f:
        for v in V:
            f.write("v %.9f %.9f %.9f\n" % (v[0], v[1], v[2]))
        for t in F:
            f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


# --------------------------------------------------------------------------
# assemblies
# --------------------------------------------------------------------------
def _spacing_for(area, n_vertices):
    return float(np.sqrt(area / float(n_vertices)))


def peg_in_hole(n_vertices=10000, peg_z=(-2.0, 4.0)):
    """C1/C2/C3 pair: peg 2x2x6 in a 6x6x4 block with a 2.1x2.1 through-hole.

    Part 0 = peg (the moving part), part 1 = block. Returned in the assembled
    configuration (peg axis = hole axis = z). ~n_vertices vertices each.
    """
    peg_area = 2 * 4.0 + 4 * 12.0
    blk_area = 4 * 24.0 + 2 * (36.0 - 2.1 * 2.1) + 4 * 2.1 * 4.0
    Vp, Fp = box_mesh((2.0, 2.0, peg_z[1] - peg_z[0]), (0.0, 0.0, 0.5 * (peg_z[0] + peg_z[1])),
                      _spacing_for(peg_area, n_vertices))
    Vb, Fb = perforated_block_mesh((6.0, 6.0, 4.0), [(0.0, 0.0)], 2.1, (0.0, 0.0, 0.0),
                                   _spacing_for(blk_area, n_vertices))
    return {"0": (Vp, Fp), "1": (Vb, Fb)}


def peg_array(nx=4, ny=5, n_vertices=10000, n_parts=20):
    """C4 assembly: a perforated plate (part 0) with pegs seated in an nx x ny
    array of holes; n_parts parts in total (plate + n_parts-1 pegs)."""
    pitch, hole, peg = 2.0, 1.1, 1.0
    plate = (pitch * nx + 1.0, pitch * ny + 1.0, 2.0)
    scale = 10.0 / max(plate)  # assembly longest side = 10 (assets/normalize.py:29-42)
    centers = [((i - (nx - 1) / 2) * pitch, (j - (ny - 1) / 2) * pitch) for j in range(ny) for i in range(nx)]
    plate_area = 2 * (plate[0] * plate[1] - len(centers) * hole * hole) + 2 * (plate[0] + plate[1]) * plate[2] \
        + len(centers) * 4 * hole * plate[2]
    Vpl, Fpl = perforated_block_mesh(plate, centers, hole, (0, 0, 0), _spacing_for(plate_area, n_vertices))
    parts = {"0": (Vpl * scale, Fpl)}
    peg_area = 2 * peg * peg + 4 * peg * 4.0
    for idx, (cx, cy) in enumerate(centers[: n_parts - 1]):
        Vp, Fp = box_mesh((peg, peg, 4.0), (cx, cy, 1.0), _spacing_for(peg_area, n_vertices))
        parts[str(idx + 1)] = (Vp * scale, Fp)
    return parts


def l_channel(h=0.25, clearance=0.1):
    """Depth-2 planning scene: a unit cube (part 0) at the blind end of an L-shaped channel
    cut into a 5x3x4 block (part 1). The cube has to slide +x to the corner of the channel
    and then leave through the vertical shaft (+z); every other axis move stalls at once."""
    w = 0.5 + clearance
    xs = _axis([-2.0, -w, 1.5 - w, 1.5 + w, 3.0], h)
    ys = _axis([-1.5, -w, w, 1.5], h)
    zs = _axis([-2.0, -w, w, 2.0], h)
    xc, yc, zc = (0.5 * (a[:-1] + a[1:]) for a in (xs, ys, zs))
    X, Y, Z = np.meshgrid(xc, yc, zc, indexing="ij")
    tunnel = (X > -w) & (X < 1.5 + w) & (np.abs(Y) < w) & (np.abs(Z) < w)
    shaft = (np.abs(X - 1.5) < w) & (np.abs(Y) < w) & (Z > -w)
    Vb, Fb = lattice_surface(xs, ys, zs, ~(tunnel | shaft))
    Vc, Fc = box_mesh((1.0, 1.0, 1.0), (0.0, 0.0, 0.0), h)
    return {"0": (Vc, Fc), "1": (Vb, Fb)}


def l_channel_with_plug(h=0.25, clearance=0.1):
    """Three-part sequence-planning scene: the L-channel block (part 1) and cube (part 0) of
    l_channel() plus a plug (part 2) sitting in the vertical shaft. The cube can only leave
    after the plug has been pulled out (+z)."""
    parts = l_channel(h, clearance)
    Vp, Fp = box_mesh((1.0, 1.0, 1.0), (1.5, 0.0, 1.4), h)
    parts["2"] = (Vp, Fp)
    return parts


def reorient(parts, R):
    """The same assembly with every vertex rotated by R (3x3)."""
    R = np.asarray(R, dtype=np.float64)
    return {k: (V @ R.T, F) for k, (V, F) in parts.items()}


def write_assembly(parts, folder):
    os.makedirs(folder, exist_ok=True)
    for pid, (V, F) in parts.items():
        write_obj(os.path.join(folder, f"{pid}.obj"), V, F)
    return folder


# --------------------------------------------------------------------------
# simulation XML (strings handed to Simulation(xml_string, asset_folder))
# --------------------------------------------------------------------------
def _arr(a):
    return " ".join(str(x) for x in a)


def planner_xml(assembly_dir, move_id, still_ids, joint_type="translational", sdf_dx=0.05, col_th=0.01,
                save_sdf=False, kn="1e6", kt="1e3", mu="0", damping="0", translation=None):
    """Same document examples/run_joint_plan.py:140-178 builds (body_type SDF)."""
    ids = [move_id, *still_ids]
    if translation is None:
        translation = {p: [0, 0, 0] for p in ids}
    sdf_args = 'load_sdf="true" save_sdf="true"' if save_sdf else ""
    s = f'''
<redmax model="assemble">
<option integrator="BDF1" timestep="1e-3" gravity="0. 0. 1e-12"/>
<render bg_color="255 255 255 255"/>

<default>
    <general_SDF_contact kn="{kn}" kt="{kt}" mu="{mu}" damping="{damping}"/>
</default>
'''
    for p in ids:
        jt = joint_type if p == move_id else "fixed"
        s += f'''
<robot>
    <link name="part{p}">
        <joint name="part{p}" type="{jt}" axis="0. 0. 0." pos="{_arr(translation[p])}" quat="1 0 0 0" frame="WORLD" damping="0"/>
        <body name="part{p}" type="SDF" filename="{assembly_dir}/{p}.obj" {sdf_args} pos="0 0 0" quat="1 0 0 0" scale="1 1 1" transform_type="OBJ_TO_JOINT" density="1" dx="{sdf_dx}" col_th="{col_th}" mu="0" rgba="0.5 0.5 0.5 1"/>
    </link>
</robot>
'''
    s += "\n<contact>\n"
    for p in still_ids:
        s += f'''
    <general_SDF_contact general_body="part{move_id}" SDF_body="part{p}"/>
    <general_SDF_contact general_body="part{p}" SDF_body="part{move_id}"/>
'''
    s += "\n</contact>\n</redmax>\n"
    return s


def freefall_xml(assembly_dir, sdf_dx=0.05, gravity=9.8, friction=0, fixed=False, ground=True):
    """Same document examples/test_joint_sim.py:24-63 builds (body_type SDF)."""
    jt = "fixed" if fixed else "free3d-exp"
    s = f'''
<redmax model="assemble">
<option integrator="BDF1" timestep="1e-3" gravity="0. 0. {-gravity}"/>
'''
    if ground:
        s += '''
<ground pos="0 0 -5" normal="0 0 1"/>'''
    s += f'''
<default>
    <ground_contact kn="1e6" kt="1e3" mu="0.8" damping="5e1"/>
    <general_SDF_contact kn="1e5" kt="1e3" mu="{friction}" damping="0"/>
</default>
'''
    for p in ["0", "1"]:
        s += f'''
<robot>
    <link name="part{p}">
        <joint name="part{p}" type="{jt}" axis="0. 0. 0." pos="0 0 0" quat="1 0 0 0" frame="WORLD" damping="0"/>
        <body name="part{p}" type="SDF" filename="{assembly_dir}/{p}.obj" pos="0 0 0" quat="1 0 0 0" scale="1 1 1" transform_type="OBJ_TO_JOINT" density="1" dx="{sdf_dx}" mu="0" rgba="0.5 0.5 0.5 1"/>
    </link>
</robot>
'''
    s += "\n<contact>\n"
    if ground:
        for p in ["0", "1"]:
            s += f'\n    <ground_contact body="part{p}"/>\n'
    s += '''
    <general_SDF_contact general_body="part0" SDF_body="part1"/>
    <general_SDF_contact general_body="part1" SDF_body="part0"/>
'''
    s += "\n</contact>\n</redmax>\n"
    return s


# --------------------------------------------------------------------------
# analytic grids and poses
# --------------------------------------------------------------------------
@dataclass
class GridSpec:
    ni: int
    nj: int
    nk: int
    min_box: np.ndarray
    max_box: np.ndarray
    dx: float
    dy: float
    dz: float
    col_th: float
    grid: np.ndarray  # flat fp64, index i + ni*(j + nj*k)  (sdfgen/array3.h:70-80)


def sphere_grid(n, length=10.0, radius_frac=0.35, col_th=0.0, dtype_round=np.float32):
    """C5 grid: SDF of a sphere of radius radius_frac*length centred in an n^3
    box of side `length`; values rounded to fp32 then widened (fresh-grid
    behaviour, Body/BodySDFObj.cpp:81-83). Origin/spacing are fp32-widened too."""
    d = float(np.float32(length / (n - 1)))
    lo = float(np.float32(-0.5 * length))
    ax = lo + d * np.arange(n, dtype=np.float64)
    r = radius_frac * length
    x2 = ax * ax
    # x fastest: index i + n*(j + n*k)
    g = np.sqrt(x2[None, None, :] + x2[None, :, None] + x2[:, None, None]) - r
    if dtype_round is not None:
        g = g.astype(dtype_round).astype(np.float64)
    mn = np.array([lo, lo, lo], dtype=np.float64)
    mx = np.array([float(ax[-1])] * 3, dtype=np.float64)
    return GridSpec(n, n, n, mn, mx, d, d, d, float(col_th), np.ascontiguousarray(g.reshape(-1)))


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)


def axis_angle_rotation(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def make_E(R, p):
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = p
    return E


def perturbed_pose(E, rng, trans=0.05, rot=0.02):
    """A pose near E: small seeded rotation (axis-angle <= rot rad) and translation."""
    ax = rng.normal(size=3)
    R = axis_angle_rotation(ax, rng.uniform(-rot, rot)) @ E[:3, :3]
    p = E[:3, 3] + rng.uniform(-trans, trans, size=3)
    return make_E(R, p)
