"""Drive the UNMODIFIED reference (oracle/_ref/ref_probe) on a synthetic scene.

TEST INFRASTRUCTURE ONLY (checker + CPU baseline): used by tests/, bench.py's
cpu_baseline leg and `bench.py --impl reference`. The reference Simulation is
constructed from the planner's XML on cheap placeholder meshes, then the
hot-path inputs it derived from them are overwritten, through public data
members, with the scene's arrays (contact points, SDF grids, col_th, force
parameters, body frames and mass properties) — so the reference's own code
evaluates exactly the inputs the CUDA path is given.
"""
from __future__ import annotations

import glob
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def available():
    return bool(glob.glob(os.path.join(REF_DIR, "ref_probe*.so")))


def _import_probe():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import ref_probe
    return ref_probe


def grid_dict(g):
    return {"ni": g.ni, "nj": g.nj, "nk": g.nk, "min_box": g.min_box, "max_box": g.max_box, "dx": g.dx, "dy": g.dy,
            "dz": g.dz, "col_th": g.col_th, "grid": g.grid}


def make_probe(bodies, kn, kt, mu, damping, joint_type="translational", inject_frames=False):
    """bodies: assemble_them_all_b200.scenes.SceneBody list, body 0 = moving part.
    -> ref_probe.Probe whose Robot holds the 2*(n-1) contact forces in planner order."""
    from assemble_them_all_b200 import synth
    ref_probe = _import_probe()
    tmp = tempfile.mkdtemp(prefix="ref_scene_")
    ids = [str(i) for i in range(len(bodies))]
    for pid in ids:
        V, F = synth.box_mesh((1.0, 1.0, 1.0), h=0.5)  # placeholder geometry, overwritten below
        synth.write_obj(os.path.join(tmp, f"{pid}.obj"), V, F)
    xml = synth.planner_xml(tmp, ids[0], ids[1:], joint_type=joint_type, sdf_dx=0.5, col_th=bodies[0].grid.col_th)
    p = ref_probe.Probe(xml, tmp)
    for b, body in enumerate(bodies):
        name = f"part{b}"
        g = body.grid
        p.set_contact_points(name, np.ascontiguousarray(body.points))
        p.set_sdf_grid(name, g.ni, g.nj, g.nk, g.min_box, g.max_box, g.dx, g.dy, g.dz, g.grid)
        p.set_col_th(name, g.col_th)
        if inject_frames:
            p.set_body_frame(name, np.asarray(body.rest_E, dtype=np.float64), np.asarray(body.inertia, dtype=np.float64),
                             float(body.mass))
    for fi in range(p.num_forces()):
        p.set_force_params(fi, kn, kt, mu, damping, 1.0)
    if inject_frames:
        p.update_robot()
    return p


def rollout(p, q0, qdot0, f_ext0, n_steps, trajectory=False, move_name="part0"):
    """reset -> set_state -> external force on part0 -> forward: what BFSPlanner does
    per action (run_joint_plan.py:543-554). -> (seconds in forward, q, qdot[, Q, QD])."""
    p.reset()
    p.set_state_init(np.asarray(q0, dtype=np.float64), np.asarray(qdot0, dtype=np.float64))
    p.reset()
    p.set_body_external_force(move_name, np.asarray(f_ext0, dtype=np.float64))
    if not trajectory:
        t = p.time_forward(int(n_steps))
        return t, p.get_q(), p.get_qdot()
    Q, QD, t = [], [], 0.0
    for _ in range(int(n_steps)):
        t += p.time_forward(1)
        Q.append(p.get_q())
        QD.append(p.get_qdot())
    return t, p.get_q(), p.get_qdot(), np.array(Q), np.array(QD)


def probe_from_assembly(parts, move_id, still_ids, joint_type="translational", sdf_dx=0.1, col_th=0.01, **xml_kw):
    """Reference Simulation built the way the planners build it: OBJ files on disk +
    the planner XML (the reference's own tinyobj loader, principal-axes frames, sdfgen)."""
    from assemble_them_all_b200 import synth
    ref_probe = _import_probe()
    tmp = tempfile.mkdtemp(prefix="ref_asm_")
    synth.write_assembly(parts, tmp)
    xml = synth.planner_xml(tmp, move_id, still_ids, joint_type=joint_type, sdf_dx=sdf_dx, col_th=col_th, **xml_kw)
    return ref_probe.Probe(xml, tmp)


def extract_scene(p):
    """Everything the contact path and the stepper consume, read back from a constructed
    reference Simulation: per body (robot order) contact points, SDF grid fields, joint
    type, E_pj_0, E_ji, inertia, mass; per force (force order) bodies and parameters."""
    names = p.body_names()
    jt = {0: "fixed", 3: "translational", 6: "free3d-exp"}
    bodies = []
    for n in names:
        bi, ji = p.body_info(n), p.joint_info(n)
        bodies.append({"name": n, "points": np.asarray(p.contact_points(n)), "grid": dict(p.sdf_grid(n)),
                       "joint_type": jt[int(ji["ndof"])], "E_pj_0": np.asarray(ji["E_pj_0"]),
                       "E_ji": np.asarray(bi["E_ji"]), "inertia": np.asarray(bi["inertia"]), "mass": float(bi["mass"]),
                       "index": int(bi["index"])})
    forces, grounds = [], []
    for fi in range(p.num_forces()):
        info = p.force_info(fi)
        if info["type"] == "ground_contact":
            grounds.append({"body": names.index(info["contact_body"]), "E_g": np.asarray(info["E_g"]), "kn": info["kn"],
                            "kt": info["kt"], "mu": info["mu"], "damping": info["damping"]})
            continue
        forces.append({"a": names.index(info["contact_body"]), "b": names.index(info["sdf_body"]), "kn": info["kn"],
                       "kt": info["kt"], "mu": info["mu"], "damping": info["damping"]})
    return {"bodies": bodies, "forces": forces, "grounds": grounds, "h": float(p.timestep()), "gravity": np.asarray(p.gravity())}
