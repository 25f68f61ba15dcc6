"""Small stepper + batch run for compute-sanitizer (initcheck / memcheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from assemble_them_all_b200 import contact, scenes  # noqa: E402

ctx = contact.Context(0)
for jt in ("translational", "free3d-exp"):
    rs = scenes.c2_planner_rollouts(n_envs=12, n_vertices=600, sdf_dx=0.1, joint_type=jt)
    rs.q0[:, 2] = np.linspace(0.0, 2.5, 12)
    sim, batch, *keep = rs.upload(ctx)
    sim.reset(rs.q0, rs.qdot0)
    for b in range(len(rs.bodies)):
        sim.set_body_external_force(b, rs.f_ext[:, b])
    sim.forward(45)
    print(jt, sim.stats(), batch.skin_stats(), flush=True)
    sim.close()
rs = scenes.c1_freefall_rollouts(n_envs=8, n_vertices=600, sdf_dx=0.1, steps=30)
sim, batch, *keep = rs.upload(ctx)
sim.reset(rs.q0, rs.qdot0)
sim.forward(30)
print("c1", sim.stats(), flush=True)
sim.close()
sc = scenes.c2_bfs_rollouts(n_envs=8, n_vertices=600, sdf_dx=0.1)
batch, *keep = sc.upload(ctx)
out = batch.eval(sc.E, sc.phi, deriv=True)
print("batch", float(np.abs(out[0]).max()), flush=True)
