#!/usr/bin/env python
"""Drive a `redmax_py` module the way the BFS planner does (examples/run_joint_plan.py:
511-570 of the reference): per action reset, set_joint_q, zero_joint_qdot,
set_body_external_force, then frame_skip x forward(1) reading the joint state after every
substep. Runs in its own process so that the stock module (oracle/_ref) and the B200
build (oracle/_ref/b200) — both named redmax_py — can be compared on identical inputs.

    python tests/drivers/redmax_rollout.py MODULE_DIR OUT.npz [--joint translational|free3d-exp]
                                           [--nverts 600] [--dx 0.1] [--steps 100]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("module_dir")
    ap.add_argument("out")
    ap.add_argument("--joint", default="translational")
    ap.add_argument("--nverts", type=int, default=600)
    ap.add_argument("--dx", type=float, default=0.1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--scene", default="planner", choices=["planner", "freefall"])
    ap.add_argument("--reps", type=int, default=1, help="repeat the rollouts and report the fastest pass (timing runs)")
    args = ap.parse_args()
    sys.path.insert(0, os.path.abspath(args.module_dir))
    import redmax_py as redmax
    from assemble_them_all_b200 import synth

    tmp = tempfile.mkdtemp(prefix="rollout_asm_")
    synth.write_assembly(synth.peg_in_hole(n_vertices=args.nverts), tmp)
    if args.scene == "freefall":
        return freefall(redmax, synth, tmp, args)
    xml = synth.planner_xml(tmp, "0", ["1"], joint_type=args.joint, sdf_dx=args.dx)
    t0 = time.perf_counter()
    sim = redmax.Simulation(xml, tmp)
    t_build = time.perf_counter() - t0
    ndof = 3 if args.joint == "translational" else 6
    dirs = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    actions = []
    for a in range(6):
        f = np.zeros(6)
        f[3:] = dirs[a] * 100.0
        actions.append(f)
    if ndof == 6:
        for a in range(6):
            f = np.zeros(6)
            f[:3] = 3.0 * dirs[a] * 100.0   # apply_action, run_joint_plan.py:270-278
            actions.append(f)
    Q = np.zeros((len(actions), args.steps, ndof))
    QD = np.zeros_like(Q)
    conv = np.zeros((len(actions), args.steps), dtype=bool)
    touched = []
    q_start = np.zeros(ndof)
    q_start[2] = 1.25
    t_best = float("inf")
    for rep in range(max(1, args.reps)):
        t_fwd = 0.0
        touched = []
        for ai, f in enumerate(actions):
            sim.reset()
            sim.set_joint_q("part0", q_start)
            sim.zero_joint_qdot("part0")
            sim.update_robot()
            sim.set_body_external_force("part0", f)
            for s in range(args.steps):
                t0 = time.perf_counter()
                sim.forward(1)
                t_fwd += time.perf_counter() - t0
                Q[ai, s] = sim.get_joint_q("part0")
                QD[ai, s] = sim.get_joint_qdot("part0")
                conv[ai, s] = sim.is_converged()
            touched.append(",".join(sim.get_contact_bodies("part0")))
        t_best = min(t_best, t_fwd)
    t_fwd = t_best
    E_final = np.asarray(sim.get_body_E0i("part0"))
    np.savez(args.out, Q=Q, QD=QD, converged=conv, touched=np.array(touched), E_final=E_final, t_forward=t_fwd,
             t_build=t_build, n_env_steps=len(actions) * args.steps)
    print(f"{os.path.basename(os.path.abspath(args.module_dir))}: {len(actions) * args.steps} env-steps in {t_fwd:.3f} s "
          f"({len(actions) * args.steps / t_fwd:.1f} env-steps/s), build {t_build:.2f} s")


def freefall(redmax, synth, tmp, args):
    """examples/test_joint_sim.py: two free parts, ground plane, gravity 9.8; plus the SDFMesh
    surface the planners use (python_interface.cpp:283-292). Exercises ForceGroundContact,
    ForceGeneralSDFContact and both sdfgen call sites of the module under test."""
    t0 = time.perf_counter()
    sim = redmax.Simulation(synth.freefall_xml(tmp, sdf_dx=args.dx, gravity=9.8, friction=0), tmp)
    t_build = time.perf_counter() - t0
    sim.reset()
    # start just above the plane so that the run covers free flight, impact and sliding
    sim.set_joint_q("part0", np.array([0.0, 0.0, -2.6, 0.03, -0.02, 0.01]))
    sim.set_joint_q("part1", np.array([0.0, 0.0, -2.9, 0.0, 0.0, 0.0]))
    sim.update_robot()
    Q = np.zeros((1, args.steps, 12))
    QD = np.zeros_like(Q)
    conv = np.zeros((1, args.steps), dtype=bool)
    t_fwd = 0.0
    for s in range(args.steps):
        t0 = time.perf_counter()
        sim.forward(1)
        t_fwd += time.perf_counter() - t0
        Q[0, s] = np.concatenate([sim.get_joint_q("part0"), sim.get_joint_q("part1")])
        QD[0, s] = np.concatenate([sim.get_joint_qdot("part0"), sim.get_joint_qdot("part1")])
        conv[0, s] = sim.is_converged()
    touched = [",".join(sorted(sim.get_contact_bodies(n))) for n in ("part0", "part1")]
    # SDFMesh: grid built by the module's own sdfgen call (Mesh.cpp:90), sampled on seeded points
    V, F = synth.peg_in_hole(n_vertices=args.nverts)["1"]
    mesh = redmax.SDFMesh(np.ascontiguousarray(V.T), np.ascontiguousarray(F.T.astype(np.int32)), args.dx)
    X = np.random.default_rng(5).uniform(-3.2, 3.2, size=(3, 4000))
    dist = np.asarray(mesh.distance(X))
    min_dist = float(mesh.min_distance(np.ascontiguousarray(X[:, :500])))
    np.savez(args.out, sdfmesh_min_distance=min_dist, Q=Q, QD=QD, converged=conv, touched=np.array(touched), E_final=np.asarray(sim.get_body_E0i("part0")),
             t_forward=t_fwd, t_build=t_build, n_env_steps=args.steps, sdfmesh_distance=dist)
    print(f"{os.path.basename(os.path.abspath(args.module_dir))}: freefall {args.steps} steps in {t_fwd:.3f} s, build {t_build:.2f} s")


if __name__ == "__main__":
    main()
