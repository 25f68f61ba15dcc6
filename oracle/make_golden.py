"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

The reference ships no golden vectors or known-answer tests for this path
(SURVEY.md §4, §8c), so the fixtures are outputs of the reference itself, run in
the build container through oracle/_ref/ref_probe (which needs /root/reference
only at BUILD time). Run from the repo root:

    make -C oracle/ref_build -j8 && python oracle/make_golden.py

Fixtures (all seeded, small):
  contact_small.npz   peg/block pair (~600 vertices each, dx 0.1): body-frame
                      contact points, fp32-lossless SDF grids, and for a set of
                      seeded poses/velocities/parameters the reference's per-point
                      d, cell index, active ids, and fm_sub/Km_sub/Dm_sub from both
                      the derivative and value-only kernels, plus raw
                      query_SDF/query_dSDF/query_ddSDF samples.
  traj_small.npz      per-step q, qdot of the reference's BDF1 integration of the
                      same pair under the planner's XML (translational and
                      free3d-exp joints, external push), for the batched stepper.
  freefall_small.npz  the free-fall configuration (examples/test_joint_sim.py:24-63: two
                      free3d-exp parts, gravity 9.8, ground plane z = -5 with mu 0.8 /
                      damping 50, SDF contact kn 1e5): the constructed scene (points,
                      grids, frames, mass properties, forces), ForceGroundContact outputs
                      (contact ids, d, fm/Km/Dm) for seeded poses and parameter sets that
                      hit the static and the dynamic Coulomb branch, and a 400-step drop
                      trajectory of both parts.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from assemble_them_all_b200 import synth  # noqa: E402

import ref_probe  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (mu, damping, kn, kt, scale, contact_scale); first rows = planner defaults
PARAM_CASES = [
    (0.0, 0.0, 1e6, 1e3, 1.0, 1.0),
    (0.0, 0.0, 1e6, 1e3, 1.0, 1.0),
    (0.0, 0.0, 1e5, 1e3, 0.5, 2.0),
    (0.5, 0.0, 1e6, 1e3, 1.0, 1.0),
    (0.0, 30.0, 1e6, 1e3, 1.0, 1.0),
    (0.5, 30.0, 1e6, 1e3, 1.0, 1.0),
    (0.9, 5.0, 1e5, 1e3, 1.0, 1.0),
    (0.05, 100.0, 1e6, 2e3, 1.0, 0.7),
]


def pack_grid(out, prefix, g):
    grid = np.asarray(g["grid"])
    g32 = grid.astype(np.float32)
    assert np.array_equal(g32.astype(np.float64), grid), "fresh grids must be fp32-lossless (BodySDFObj.cpp:81-83)"
    out[prefix + "grid_f32"] = g32
    out[prefix + "dims"] = np.array([g["ni"], g["nj"], g["nk"]], dtype=np.int32)
    out[prefix + "min_box"] = np.asarray(g["min_box"])
    out[prefix + "max_box"] = np.asarray(g["max_box"])
    out[prefix + "spacing"] = np.array([g["dx"], g["dy"], g["dz"]])
    out[prefix + "col_th"] = np.array(g["col_th"])


def make_contact_small():
    rng = np.random.default_rng(synth.MESH_SEED)
    tmp = tempfile.mkdtemp(prefix="golden_asm_")
    parts = synth.peg_in_hole(n_vertices=600)
    synth.write_assembly(parts, tmp)
    xml = synth.planner_xml(tmp, "0", ["1"], sdf_dx=0.1)
    p = ref_probe.Probe(xml, tmp)
    names = p.body_names()
    out = {"names": np.array(names), "n_forces": np.array(p.num_forces()), "n_cases": np.array(len(PARAM_CASES)),
           "ndof_m": np.array(p.ndof_m())}
    E0 = {}
    for n in names:
        out[f"{n}_points"] = p.contact_points(n)
        pack_grid(out, f"{n}_", p.sdf_grid(n))
        E0[n] = p.body_info(n)["E_0i"]
        out[f"{n}_E0"] = E0[n]
        out[f"{n}_index"] = np.array(p.body_info(n)["index"])
    for fi in range(p.num_forces()):
        info = p.force_info(fi)
        out[f"f{fi}_bodies"] = np.array([info["contact_body"], info["sdf_body"]])

    # raw SDF samples (inside, near faces, outside)
    for n in names:
        g = p.sdf_grid(n)
        lo, hi = np.asarray(g["min_box"]), np.asarray(g["max_box"])
        X = rng.uniform(lo - 0.3, hi + 0.3, size=(3000, 3))
        # exact lattice / box-face hits exercise floor() and the closed-interval test
        ijk = rng.integers(0, [g["ni"], g["nj"], g["nk"]], size=(200, 3))
        X[:200] = lo + ijk * np.array([g["dx"], g["dy"], g["dz"]])
        X[200] = lo
        X[201] = hi
        out[f"{n}_qX"] = X
        out[f"{n}_q_sdf"] = p.query_sdf(n, X, False)
        out[f"{n}_q_sdf_acc"] = p.query_sdf(n, X, True)
        out[f"{n}_q_dsdf"] = p.query_dsdf(n, X)
        out[f"{n}_q_ddsdf"] = p.query_ddsdf(n, X[:1000])
        out[f"{n}_q_cell"] = p.grid_location(n, X).astype(np.int32)

    for c, (mu, damp, kn, kt, scale, cscale) in enumerate(PARAM_CASES):
        out[f"c{c}_params"] = np.array([kn, kt, mu, damp, scale, cscale])
        for n in names:
            big = c in (1,)  # one deep-penetration case
            E = synth.perturbed_pose(E0[n], rng, trans=0.2 if big else 0.08, rot=0.08 if big else 0.03)
            phi = rng.normal(size=6) * (0.0 if c == 0 else 0.5)
            p.set_body_pose(n, E, phi)
            p.set_contact_scale(n, cscale)
            out[f"c{c}_{n}_E"] = E
            out[f"c{c}_{n}_phi"] = phi
        p.clear_all_contact_bodies()
        for fi in range(p.num_forces()):
            p.set_force_params(fi, kn, kt, mu, damp, scale)
            det = p.detect(fi)
            out[f"c{c}_f{fi}_d"] = det["d"]
            out[f"c{c}_f{fi}_cell"] = det["cell"].astype(np.int32)
            out[f"c{c}_f{fi}_active"] = np.array(det["active_ids"], dtype=np.int32)
            fm, Km, Dm = p.force_deriv_sub(fi)
            out[f"c{c}_f{fi}_fm"] = fm
            out[f"c{c}_f{fi}_Km"] = Km
            out[f"c{c}_f{fi}_Dm"] = Dm
            out[f"c{c}_f{fi}_fm_value"] = p.force_sub(fi)
            fm_sys, Km_sys, Dm_sys = p.force_deriv_public(fi)
            out[f"c{c}_f{fi}_fm_sys"] = fm_sys
            out[f"c{c}_f{fi}_Km_sys"] = Km_sys
            out[f"c{c}_f{fi}_Dm_sys"] = Dm_sys
            if c in (1, 5):
                ids, FM, KM, DM = p.force_deriv_per_contact(fi)
                out[f"c{c}_f{fi}_pc_ids"] = np.array(ids, dtype=np.int32)
                out[f"c{c}_f{fi}_pc_fm"] = FM
                out[f"c{c}_f{fi}_pc_Km"] = KM
                out[f"c{c}_f{fi}_pc_Dm"] = DM
        for n in names:
            out[f"c{c}_{n}_touched"] = np.array(p.get_contact_bodies(n))
    path = os.path.join(GOLDEN, "contact_small.npz")
    np.savez_compressed(path, **out)
    return path


def make_traj_small(steps=120):
    tmp = tempfile.mkdtemp(prefix="golden_traj_")
    parts = synth.peg_in_hole(n_vertices=600)
    synth.write_assembly(parts, tmp)
    out = {"steps": np.array(steps)}
    for jt, force in (("translational", [0, 0, 0, 60.0, 80.0, 0.0]),
                      ("free3d-exp", [30.0, -60.0, 15.0, 60.0, 80.0, 0.0])):
        key = jt.replace("-", "_")
        xml = synth.planner_xml(tmp, "0", ["1"], joint_type=jt, sdf_dx=0.1)
        p = ref_probe.Probe(xml, tmp)
        p.reset()
        p.set_body_external_force("part0", np.array(force, dtype=np.float64))
        Q, QD, CONV = [p.get_q()], [p.get_qdot()], []
        for _ in range(steps):
            p.forward(1)
            Q.append(p.get_q())
            QD.append(p.get_qdot())
            CONV.append(p.is_converged())
        out[f"{key}_force"] = np.array(force, dtype=np.float64)
        out[f"{key}_q"] = np.array(Q)
        out[f"{key}_qdot"] = np.array(QD)
        out[f"{key}_converged"] = np.array(CONV)
        out[f"{key}_touched"] = np.array(p.get_contact_bodies("part0"))
        out[f"{key}_h"] = np.array(p.timestep())
        for n in p.body_names():
            bi = p.body_info(n)
            out[f"{key}_{n}_mass"] = np.array(bi["mass"])
            out[f"{key}_{n}_inertia"] = np.asarray(bi["inertia"])
            out[f"{key}_{n}_E_ij"] = np.asarray(bi["E_ij"])
            out[f"{key}_{n}_E_final"] = np.asarray(bi["E_0i"])
    path = os.path.join(GOLDEN, "traj_small.npz")
    np.savez_compressed(path, **out)
    return path


GROUND_CASES = [  # (kn, kt, mu, damping)
    (1e6, 1e3, 0.8, 5e1),   # test_joint_sim.py:36
    (1e6, 1e3, 0.0, 0.0),
    (1e6, 1e3, 0.02, 5e1),  # dynamic branch for most contacts
    (1e5, 2e3, 0.4, 0.0),
]


def make_freefall_small(steps=400):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import ref_scene
    rng = np.random.default_rng(synth.MESH_SEED + 5)
    tmp = tempfile.mkdtemp(prefix="golden_fall_")
    synth.write_assembly(synth.peg_in_hole(n_vertices=600), tmp)
    p = ref_probe.Probe(synth.freefall_xml(tmp, sdf_dx=0.1, gravity=9.8, friction=0), tmp)
    sc = ref_scene.extract_scene(p)
    out = {"names": np.array(p.body_names()), "h": np.array(sc["h"]), "gravity": sc["gravity"],
           "n_forces": np.array(len(sc["forces"])), "n_grounds": np.array(len(sc["grounds"])),
           "n_cases": np.array(len(GROUND_CASES)), "steps": np.array(steps)}
    for b in sc["bodies"]:
        n = b["name"]
        out[f"{n}_points"] = b["points"]
        pack_grid(out, f"{n}_", b["grid"])
        for k in ("E_pj_0", "E_ji", "inertia"):
            out[f"{n}_{k}"] = np.asarray(b[k])
        out[f"{n}_mass"] = np.array(b["mass"])
        out[f"{n}_joint_type"] = np.array(b["joint_type"])
    for i, f in enumerate(sc["forces"]):
        out[f"f{i}"] = np.array([f["a"], f["b"], f["kn"], f["kt"], f["mu"], f["damping"]], dtype=np.float64)
    for i, g in enumerate(sc["grounds"]):
        out[f"g{i}_body"] = np.array(g["body"])
        out[f"g{i}_E_g"] = g["E_g"]
        out[f"g{i}_params"] = np.array([g["kn"], g["kt"], g["mu"], g["damping"]])
    # which entries of Robot::_forces are the ground contacts
    gfi = [fi for fi in range(p.num_forces()) if p.force_info(fi)["type"] == "ground_contact"]
    n_static_dynamic = [0, 0]
    for c, (kn, kt, mu, damp) in enumerate(GROUND_CASES):
        out[f"c{c}_params"] = np.array([kn, kt, mu, damp])
        q = np.zeros(12)
        qd = rng.normal(size=12) * (0.0 if c == 1 else 0.6)
        q[2] = -3.02 - 0.01 * c          # peg: lowest face slightly below the plane
        q[8] = -3.01 - 0.01 * c          # block
        q[3:6] = rng.normal(size=3) * 0.04
        q[9:12] = rng.normal(size=3) * 0.02
        p.set_state(q, qd)
        p.update_robot()
        out[f"c{c}_q"], out[f"c{c}_qdot"] = q, qd
        for gi, fi in enumerate(gfi):
            p.set_force_params(fi, kn, kt, mu, damp, 1.0)
            name = p.force_info(fi)["contact_body"]
            bi = p.body_info(name)
            idx = bi["index"]
            out[f"c{c}_g{gi}_E"] = np.asarray(bi["E_0i"])
            out[f"c{c}_g{gi}_phi"] = np.asarray(bi["phi"])
            det = p.ground_detect(fi)
            out[f"c{c}_g{gi}_active"] = np.array(det["active_ids"], dtype=np.int32)
            out[f"c{c}_g{gi}_d"] = np.array(det["d"], dtype=np.float64)
            fm, Km, Dm = p.force_deriv_public(fi)
            out[f"c{c}_g{gi}_fm"] = fm[idx:idx + 6]
            out[f"c{c}_g{gi}_Km"] = Km[idx:idx + 6, idx:idx + 6]
            out[f"c{c}_g{gi}_Dm"] = Dm[idx:idx + 6, idx:idx + 6]
            out[f"c{c}_g{gi}_fm_value"] = p.force_public(fi)[idx:idx + 6]
            assert len(det["active_ids"]) > 0
    # restore the XML parameters and drop both parts onto the plane
    for gi, fi in enumerate(gfi):
        g = sc["grounds"][gi]
        p.set_force_params(fi, g["kn"], g["kt"], g["mu"], g["damping"], 1.0)
    q0 = np.zeros(12)
    q0[2], q0[8] = -2.5, -2.9
    q0[3:6] = [0.03, -0.02, 0.01]
    qd0 = np.zeros(12)
    qd0[0], qd0[7] = 0.4, -0.3
    p.reset()
    p.set_state(q0, qd0)
    Q, QD, CONV = [p.get_q()], [p.get_qdot()], []
    for _ in range(steps):
        p.forward(1)
        Q.append(p.get_q())
        QD.append(p.get_qdot())
        CONV.append(p.is_converged())
    out["traj_q"], out["traj_qdot"], out["traj_converged"] = np.array(Q), np.array(QD), np.array(CONV)
    for n in p.body_names():
        out[f"traj_{n}_touched"] = np.array(p.get_contact_bodies(n))
        out[f"traj_{n}_E_final"] = np.asarray(p.body_info(n)["E_0i"])
    path = os.path.join(GOLDEN, "freefall_small.npz")
    np.savez_compressed(path, **out)
    return path


def make_sdfgen_small():
    """Grids of the reference's own sdfgen::make_level_set3 (called directly) and of a constructed
    BodySDFObj, for a convex part and a concave block with an internal L-shaped channel."""
    from oracle import ref_scene
    out = {}
    parts = synth.l_channel(h=0.3)
    p = ref_scene.probe_from_assembly(parts, "0", ["1"], "translational", sdf_dx=0.15)
    names = p.body_names()
    out["names"] = np.array(names)
    out["dx_request"] = np.array(float(np.float32(0.15)))   # the XML scalar goes through float (Simulation_Constructor.cpp:566)
    for n in names:
        V = np.array(p.get_body_vertices(n)).T.copy()
        F = np.array(p.get_body_faces(n)).T.copy().astype(np.uint32)
        g = p.sdf_grid(n)
        out[f"{n}_V"] = V
        out[f"{n}_F"] = F
        out[f"{n}_dims"] = np.array([g["ni"], g["nj"], g["nk"]], dtype=np.int32)
        out[f"{n}_min_box"] = np.asarray(g["min_box"])
        out[f"{n}_max_box"] = np.asarray(g["max_box"])
        out[f"{n}_spacing"] = np.array([g["dx"], g["dy"], g["dz"]])
        grid = np.asarray(g["grid"])
        assert np.array_equal(grid.astype(np.float32).astype(np.float64), grid)
        out[f"{n}_grid_f32"] = grid.astype(np.float32)            # BodySDFObj::precompute_SDF, exact_band = 1
        sp = np.array([g["dx"], g["dy"], g["dz"]]).astype(np.float32)
        phi2 = ref_probe.make_level_set3(F, V.astype(np.float32), np.asarray(g["min_box"]).astype(np.float32), sp[0], sp[1],
                                         sp[2], g["ni"], g["nj"], g["nk"], 2)
        out[f"{n}_grid_band2_f32"] = np.asarray(phi2)
    path = os.path.join(GOLDEN, "sdfgen_small.npz")
    np.savez_compressed(path, **out)
    return path


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    fns = {"contact": make_contact_small, "traj": make_traj_small, "freefall": make_freefall_small,
           "sdfgen": make_sdfgen_small}
    sel = [a for a in sys.argv[1:] if a in fns] or list(fns)
    for fn in (fns[k] for k in sel):
        path = fn()
        print(path, "%.1f KB" % (os.path.getsize(path) / 1024))
