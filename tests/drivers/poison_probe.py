"""Tiny contact batch + stepper run that prints a digest of its results. tests/test_contact_gpu.py runs it
twice, with and without SDFC_POISON=1 (fresh device allocations filled with 0xFF bytes): equal digests mean
the library never reads device memory it did not write."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from assemble_them_all_b200 import contact, scenes  # noqa: E402


def main():
    h = hashlib.sha256()
    ctx = contact.Context(0)
    sc = scenes.c2_bfs_rollouts(n_envs=6, n_vertices=600, sdf_dx=0.1)
    batch, *keep = sc.upload(ctx)
    for deriv in (True, False):
        for a in batch.eval(sc.E, sc.phi, deriv=deriv):
            h.update(np.ascontiguousarray(a).tobytes())
    for jt in ("translational", "free3d-exp"):
        rs = scenes.c2_planner_rollouts(n_envs=12, n_vertices=600, sdf_dx=0.1, joint_type=jt)
        rs.q0[:, 2] = np.linspace(0.0, 2.5, 12)
        sim, *keep2 = rs.upload(ctx)
        sim.reset(rs.q0, rs.qdot0)
        for b in range(len(rs.bodies)):
            sim.set_body_external_force(b, rs.f_ext[:, b])
        Q, QD = sim.forward(40, trajectory=True)
        h.update(Q.tobytes()); h.update(QD.tobytes()); h.update(sim.get_contact_history().tobytes())
        sim.close()
    rs = scenes.c1_freefall_rollouts(n_envs=6, n_vertices=600, sdf_dx=0.1, steps=25)
    sim, *keep3 = rs.upload(ctx)
    sim.reset(rs.q0, rs.qdot0)
    Q, QD = sim.forward(25, trajectory=True)
    h.update(Q.tobytes()); h.update(sim.get_ground_history().tobytes())
    assert np.isfinite(Q).all()
    sim.close()
    print("DIGEST", h.hexdigest())


if __name__ == "__main__":
    main()
