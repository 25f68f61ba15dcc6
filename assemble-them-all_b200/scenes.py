"""Benchmark / parity workloads expressed as hot-path inputs.

A ContactScene holds exactly what the contact path consumes — per body: contact
points in the body frame and a voxel SDF in the body frame; per pair: (contact
body, SDF body, parameters); per environment: body poses and velocities — so the
CUDA arm and the reference arm (which injects the same arrays into the
unmodified reference through oracle/_ref/ref_probe) evaluate identical inputs.

Datasets are not available offline: meshes come from synth.py and the SDF grids
are ANALYTIC signed distances of the same shapes, sampled on a grid laid out by
the reference's own rule (Body/BodySDFObj.cpp:59-70: spacing min(dx, bbox/res),
one cell of padding, fp32 origin) and rounded to fp32 then widened
(BodySDFObj.cpp:81-83). Body frames are the mesh frames shifted to the centroid
of the bounding box (identity rotation).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import synth


def _box_sdf(p, half):
    q = np.abs(p) - half
    outside = np.linalg.norm(np.maximum(q, 0.0), axis=-1)
    inside = np.minimum(np.max(q, axis=-1), 0.0)
    return outside + inside


def _grid_layout(vmin, vmax, dx, res=20):
    """Grid geometry as BodySDFObj::precompute_SDF lays it out (fp32 arithmetic where the reference uses it)."""
    bbox = vmax - vmin
    d = np.minimum(dx, bbox / res)                        # _dx, _dy, _dz (double)
    unit = d.astype(np.float32)
    mn = vmin.astype(np.float32) - np.float32(1) * unit   # padding = 1
    mx = vmax.astype(np.float32) + np.float32(1) * unit
    sizes = ((mx - mn + (0.5 * d).astype(np.float32)) / d.astype(np.float32)).astype(np.uint32) + 1
    return d, mn.astype(np.float64), mx.astype(np.float64), sizes.astype(int)


def analytic_grid(sdf_fn, vmin, vmax, dx, col_th, res=20):
    d, mn, mx, n = _grid_layout(np.asarray(vmin, float), np.asarray(vmax, float), dx, res)
    df = d.astype(np.float32).astype(np.float64)  # sdfgen steps in float
    ax = [mn[a] + df[a] * np.arange(n[a]) for a in range(3)]
    vals = np.empty((n[2], n[1], n[0]), dtype=np.float64)
    for k in range(n[2]):  # slab by slab keeps memory small; layout index i + ni*(j + nj*k)
        P = np.stack(np.meshgrid(ax[0], ax[1], [ax[2][k]], indexing="xy"), axis=-1).reshape(-1, 3)
        vals[k] = sdf_fn(P).reshape(n[1], n[0])
    vals = vals.astype(np.float32).astype(np.float64)
    return synth.GridSpec(int(n[0]), int(n[1]), int(n[2]), mn, mx, float(d[0]), float(d[1]), float(d[2]), float(col_th),
                          np.ascontiguousarray(vals.reshape(-1)))


@dataclass
class SceneBody:
    name: str
    points: np.ndarray       # (n,3) body frame
    grid: synth.GridSpec     # body frame
    rest_E: np.ndarray       # (4,4) pose in the assembled configuration
    mass: float = 1.0        # Body::_mass (density 1, run_joint_plan.py:162)
    inertia: np.ndarray = None   # Body::_Inertia: principal moments about the body frame (3), mass (3)
    joint_type: str = "fixed"

    def desc(self):
        """Stepper descriptor: the joint sits at the world origin (pos 0, frame WORLD,
        run_joint_plan.py:161), so E_pj_0 = I and E_ji = rest pose of the body frame."""
        return {"joint_type": self.joint_type, "E_pj_0": np.eye(4), "E_ji": self.rest_E, "inertia": self.inertia,
                "mass": self.mass}


def _box_inertia(size, density=1.0):
    a, b, c = size
    m = density * a * b * c
    return m, m / 12.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])


@dataclass
class ContactScene:
    bodies: list
    pairs: list              # (contact body index, SDF body index)
    kn: float = 1e6
    kt: float = 1e3
    mu: float = 0.0
    damping: float = 0.0
    E: np.ndarray = field(default=None)    # (n_envs, n_bodies, 4, 4)
    phi: np.ndarray = field(default=None)  # (n_envs, n_bodies, 6)
    label: str = ""

    @property
    def n_envs(self):
        return len(self.E)

    @property
    def queries_per_eval(self):
        return self.n_envs * sum(len(self.bodies[a].points) for a, _ in self.pairs)

    def upload(self, ctx, morton_sort=True):
        """-> (ContactBatch, [pairs], [sdf], [points]) on the given context."""
        from . import contact
        sdf = [contact.SdfGrid.from_spec(ctx, b.grid) for b in self.bodies]
        pts = [contact.ContactPoints(ctx, b.points, morton_sort=morton_sort) for b in self.bodies]
        pairs = [contact.ForceGeneralSDFContact(ctx, pts[a], sdf[b], self.kn, self.kt, self.mu, self.damping,
                                                index_contact=6 * a, index_sdf=6 * b) for a, b in self.pairs]
        batch = contact.ContactBatch(ctx, self.n_envs, len(self.bodies), pairs, [a for a, _ in self.pairs],
                                     [b for _, b in self.pairs])
        return batch, pairs, sdf, pts


def peg_hole_bodies(n_vertices=10000, sdf_dx=0.05, col_th=0.01):
    """The C1/C2/C3 pair: peg 2x2x6 (body 0, moving) in a 6x6x4 block with a 2.1 square through-hole (body 1)."""
    parts = synth.peg_in_hole(n_vertices=n_vertices)
    col_th = float(np.float32(col_th))  # XML scalars pass through float (Simulation_Constructor.cpp:566-574)
    sdf_dx = float(np.float32(sdf_dx))
    out = []
    half_peg = np.array([1.0, 1.0, 3.0])
    half_blk = np.array([3.0, 3.0, 2.0])

    def sd_peg(p):
        return _box_sdf(p, half_peg)

    def sd_block(p):
        hole = _box_sdf(p[:, :2], np.array([1.05, 1.05]))
        return np.maximum(_box_sdf(p, half_blk), -hole)

    m_peg, I_peg = _box_inertia(2 * half_peg)
    m_blk, I_blk = _box_inertia(2 * half_blk)
    m_hole, I_hole = _box_inertia((2.1, 2.1, 4.0))  # concentric with the block: moments subtract
    props = {"0": (m_peg, I_peg), "1": (m_blk - m_hole, I_blk - I_hole)}
    for pid, fn in (("0", sd_peg), ("1", sd_block)):
        V, _ = parts[pid]
        c = 0.5 * (V.min(0) + V.max(0))  # == centre of mass for both shapes
        Vb = V - c
        grid = analytic_grid(fn, Vb.min(0), Vb.max(0), sdf_dx, col_th)
        m, I3 = props[pid]
        out.append(SceneBody(f"part{pid}", np.ascontiguousarray(Vb), grid, synth.make_E(np.eye(3), c), float(m),
                             np.concatenate([I3, [m, m, m]])))
    return out


def c2_bfs_rollouts(n_envs=1024, n_vertices=10000, seed=synth.MESH_SEED, sdf_dx=0.05, col_th=0.01):
    """C2: states of 1024 translational BFS rollouts mid-push (run_joint_plan.py:511-570).

    Environment e applies axis action e % 6 (+x, -x, +y, -y, +z, -z) to the peg
    from a start offset along the hole axis uniform in [0, 3]; lateral pushes have
    driven the peg 0.042-0.07 into the hole wall band (clearance 0.05, col_th
    0.01), axial pushes slide freely. Planner parameters: kn 1e6, kt 1e3, mu 0,
    damping 0 (run_joint_plan.py:153)."""
    rng = np.random.default_rng(seed)
    bodies = peg_hole_bodies(n_vertices, sdf_dx, col_th)
    dirs = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    E = np.tile(np.eye(4), (n_envs, 2, 1, 1))
    phi = np.zeros((n_envs, 2, 6))
    for e in range(n_envs):
        a = dirs[e % 6]
        z0 = rng.uniform(0.0, 3.0)
        push = rng.uniform(0.042, 0.07) if abs(a[2]) < 0.5 else rng.uniform(0.0, 0.5)
        E[e, 0] = bodies[0].rest_E
        E[e, 0, :3, 3] += a * push + np.array([0.0, 0.0, z0])
        E[e, 1] = bodies[1].rest_E
        phi[e, 0, 3:] = a * rng.uniform(0.05, 0.3)  # body frame == world axes (identity rotation)
    return ContactScene(bodies, [(0, 1), (1, 0)], 1e6, 1e3, 0.0, 0.0, E, phi,
                        label=f"C2 two-part BFS translational, {n_envs} rollouts, 2x~{n_vertices} contact pts, dx {sdf_dx}")


@dataclass
class RolloutScene:
    """Start states of batched planner rollouts for the stepper (sdfc_sim_* / Simulation::forward)."""
    bodies: list
    pairs: list
    q0: np.ndarray        # (n_envs, ndof_r)
    qdot0: np.ndarray
    f_ext: np.ndarray     # (n_envs, n_bodies, 6): set_body_external_force (run_joint_plan.py:270-278)
    h: float
    gravity: np.ndarray
    kn: float = 1e6
    kt: float = 1e3
    mu: float = 0.0
    damping: float = 0.0
    frame_skip: int = 100
    label: str = ""
    grounds: list = field(default_factory=list)   # {"body", "E_g", "kn", "kt", "mu", "damping"} (ForceGroundContact)

    @property
    def n_envs(self):
        return len(self.q0)

    @property
    def queries_per_eval(self):
        return self.n_envs * sum(len(self.bodies[a].points) for a, _ in self.pairs)

    def upload(self, ctx, shared=None):
        """shared: optional {body name: (SdfGrid, ContactPoints)} of this context, filled on demand, so that
        several scenes over the same parts (C4 candidates) keep one device copy of every grid / point set."""
        from . import contact
        if shared is None:
            shared = {}
        for b in self.bodies:
            if b.name not in shared:
                shared[b.name] = (contact.SdfGrid.from_spec(ctx, b.grid), contact.ContactPoints(ctx, b.points))
        sdf = [shared[b.name][0] for b in self.bodies]
        pts = [shared[b.name][1] for b in self.bodies]
        pairs = [contact.ForceGeneralSDFContact(ctx, pts[a], sdf[b], self.kn, self.kt, self.mu, self.damping,
                                                index_contact=6 * a, index_sdf=6 * b) for a, b in self.pairs]
        batch = contact.ContactBatch(ctx, self.n_envs, len(self.bodies), pairs, [a for a, _ in self.pairs],
                                     [b for _, b in self.pairs])
        gforces = [contact.ForceGroundContact(ctx, pts[g["body"]], g["E_g"], g["kn"], g["kt"], g["mu"], g["damping"],
                                              6 * g["body"]) for g in self.grounds]
        if gforces:
            batch.set_grounds(gforces, [g["body"] for g in self.grounds])
        sim = contact.BatchedSimulation(batch, [b.desc() for b in self.bodies], self.gravity, self.h)
        sim._grounds = gforces
        return sim, batch, pairs, sdf, pts


def c2_planner_rollouts(n_envs=1024, n_vertices=10000, seed=synth.MESH_SEED, sdf_dx=0.05, col_th=0.01,
                        joint_type="translational", force_mag=100.0):
    """C2 as the BFS planner drives it (run_joint_plan.py:511-570): environment e starts
    at rest with the peg offset z0 in [0, 3] along the hole axis (a BFS node), applies
    axis action e % 6 (e % 12 with rotation) with force_mag 100 and integrates frame_skip = 100 BDF1
    steps of h = float(1e-3), gravity (0, 0, 1e-12) (run_joint_plan.py:149)."""
    rng = np.random.default_rng(seed)
    bodies = peg_hole_bodies(n_vertices, sdf_dx, col_th)
    bodies[0].joint_type = joint_type
    nr = 3 if joint_type == "translational" else 6
    dirs = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    q0 = np.zeros((n_envs, nr))
    f = np.zeros((n_envs, 2, 6))
    n_act = 6 if nr == 3 else 12
    for e in range(n_envs):
        q0[e, 2] = rng.uniform(0.0, 3.0)
        a = e % n_act
        if a < 6:
            f[e, 0, 3:] = dirs[a] * force_mag          # translation action
        else:
            f[e, 0, :3] = 3.0 * dirs[a - 6] * force_mag  # rotation action (run_joint_plan.py:275)
    return RolloutScene(bodies, [(0, 1), (1, 0)], q0, np.zeros((n_envs, nr)), f, float(np.float32(1e-3)),
                        np.array([0.0, 0.0, 1e-12]),
                        label=f"C2 two-part BFS {joint_type}, {n_envs} rollouts x frame_skip 100, 2x~{n_vertices} contact pts, dx {sdf_dx}")


def c1_freefall_rollouts(n_envs=1024, n_vertices=10000, seed=synth.MESH_SEED, sdf_dx=0.05, col_th=0.01, steps=100):
    """C1 batched: examples/test_joint_sim.py:24-63 (both parts free3d-exp, gravity 9.8, ground plane
    z = -5 with kn 1e6 / kt 1e3 / mu 0.8 / damping 50, SDF contact kn 1e5) for n_envs drops that start
    a few centimetres above the plane with seeded tilts and lateral velocities, so that the
    window covers free flight, impact and frictional sliding."""
    rng = np.random.default_rng(seed)
    bodies = peg_hole_bodies(n_vertices, sdf_dx, col_th)
    for b in bodies:
        b.joint_type = "free3d-exp"
    q0 = np.zeros((n_envs, 12))
    qd0 = np.zeros((n_envs, 12))
    q0[:, 2] = -2.0 + rng.uniform(-0.9, -0.6, n_envs)       # peg: inside the hole, lowest face 0.1 .. 0.4 above the plane
    q0[:, 8] = -3.0 + rng.uniform(0.005, 0.05, n_envs)      # block: 0.5 .. 5 cm above the plane
    q0[:, 3:6] = rng.normal(scale=0.01, size=(n_envs, 3))
    q0[:, 9:12] = rng.normal(scale=0.005, size=(n_envs, 3))
    qd0[:, 0:2] = rng.normal(scale=0.2, size=(n_envs, 2))
    qd0[:, 6:8] = rng.normal(scale=0.2, size=(n_envs, 2))
    E_g = np.eye(4)
    E_g[2, 3] = -5.0
    grounds = [{"body": b, "E_g": E_g, "kn": 1e6, "kt": 1e3, "mu": float(np.float32(0.8)), "damping": 5e1} for b in (0, 1)]
    return RolloutScene(bodies, [(0, 1), (1, 0)], q0, qd0, np.zeros((n_envs, 2, 6)), float(np.float32(1e-3)),
                        np.array([0.0, 0.0, -9.8]), kn=1e5, kt=1e3, mu=0.0, damping=0.0, frame_skip=steps, grounds=grounds,
                        label=f"C1 free fall, {n_envs} drops x {steps} steps, 2x~{n_vertices} contact pts, ground + SDF contact, gravity 9.8")


def c4_bodies(nx=4, ny=5, n_parts=20, n_vertices=10000, sdf_dx=0.05, col_th=0.01):
    """The C4 assembly (synth.peg_array): a perforated plate (part 0) and n_parts - 1 pegs seated in its
    square through-holes. Body frames are the axis-aligned frames at the centres of mass (all parts are
    symmetric about them, so the products of inertia vanish); grids are analytic box / plate distances
    laid out and rounded to fp32 as BodySDFObj::precompute_SDF does (BodySDFObj.cpp:59-83)."""
    parts = synth.peg_array(nx=nx, ny=ny, n_vertices=n_vertices, n_parts=n_parts)
    col_th = float(np.float32(col_th))
    sdf_dx = float(np.float32(sdf_dx))
    pitch, hole, peg = 2.0, 1.1, 1.0
    plate = np.array([pitch * nx + 1.0, pitch * ny + 1.0, 2.0])
    s = 10.0 / plate.max()
    centers = np.array([((i - (nx - 1) / 2) * pitch, (j - (ny - 1) / 2) * pitch) for j in range(ny) for i in range(nx)]) * s
    half_plate, half_hole, half_peg = 0.5 * plate * s, 0.5 * hole * s * np.ones(2), 0.5 * s * np.array([peg, peg, 4.0])

    def sd_plate(p):
        holes = np.min(np.stack([_box_sdf(p[:, :2] - c, half_hole) for c in centers]), axis=0)
        return np.maximum(_box_sdf(p, half_plate), -holes)

    def sd_peg(p):
        return _box_sdf(p, half_peg)

    m_pl, I_pl = _box_inertia(2 * half_plate)
    m_h, I_h = _box_inertia((hole * s, hole * s, plate[2] * s))
    for cx, cy in centers:   # parallel-axis theorem, hole by hole
        I_pl = I_pl - (I_h + m_h * np.array([cy * cy, cx * cx, cx * cx + cy * cy]))
        m_pl -= m_h
    m_pg, I_pg = _box_inertia(2 * half_peg)
    out = []
    for pid in sorted(parts, key=int):
        V, _ = parts[pid]
        c = 0.5 * (V.min(0) + V.max(0))
        Vb = V - c
        fn, m, I3 = (sd_plate, m_pl, I_pl) if pid == "0" else (sd_peg, m_pg, I_pg)
        grid = analytic_grid(fn, Vb.min(0), Vb.max(0), sdf_dx, col_th)
        out.append(SceneBody(f"part{pid}", np.ascontiguousarray(Vb), grid, synth.make_E(np.eye(3), c), float(m),
                             np.concatenate([I3, [m, m, m]])))
    return out


def c4_candidate_rollouts(bodies, move, force_mag=100.0, joint_type="translational"):
    """C4: the six axis-direction rollouts of ONE candidate removal (run_multi_plan.py:84-126 tries every
    remaining part with a BFS plan whose first stage is these actions). Body 0 of the returned scene is the
    candidate `move` (index into `bodies`), the others stay fixed in id order; 2 contact forces per still
    part in the planner XML's order (run_joint_plan.py:169-173)."""
    order = [move] + [i for i in range(len(bodies)) if i != move]
    bs = []
    for k, i in enumerate(order):
        b = bodies[i]
        bs.append(SceneBody(b.name, b.points, b.grid, b.rest_E, b.mass, b.inertia, joint_type if k == 0 else "fixed"))
    pairs = []
    for k in range(1, len(bs)):
        pairs += [(0, k), (k, 0)]
    dirs = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    nr = 3 if joint_type == "translational" else 6
    f = np.zeros((6, len(bs), 6))
    f[:, 0, 3:] = dirs * force_mag
    return RolloutScene(bs, pairs, np.zeros((6, nr)), np.zeros((6, nr)), f, float(np.float32(1e-3)),
                        np.array([0.0, 0.0, 1e-12]),
                        label=f"C4 {len(bs)}-part plate + pegs, candidate {bodies[move].name}, 6 axis rollouts x frame_skip 100")


def sweep_scene(n_points, grid_n, active_frac=0.05, seed=synth.SWEEP_SEED):
    """C5: n_points uniform in an n^3 analytic sphere grid; col_th picked for the active fraction."""
    rng = np.random.default_rng(seed)
    spec = synth.sphere_grid(grid_n)
    lo, hi = spec.min_box, spec.max_box
    pts = rng.uniform(lo * 0.999, hi * 0.999, size=(n_points, 3))
    r = np.linalg.norm(pts[: min(n_points, 200000)], axis=1) - 3.5
    spec.col_th = float(-np.quantile(r, active_frac))
    return spec, pts
