"""CPU: the numpy stepper oracle against the UNMODIFIED reference's own
Simulation::forward (oracle/_ref/ref_probe) on the synthetic planner rollouts that
bench.py times (scenes.c2_planner_rollouts, small mesh). Skipped when oracle/_ref
has not been built (it is built by __graft_entry__.build() in the build container
and shipped to the GPU box)."""
import numpy as np
import pytest

from tests.helpers import assert_state_close


def _oracle_for(sc, e):
    from oracle import sdfc_oracle as O, sim_oracle as S
    O.build()
    jt = {"fixed": S.FIXED, "translational": S.TRANSLATIONAL, "free3d-exp": S.FREE3D_EXP}
    grids = [O.Grid.from_spec(b.grid) for b in sc.bodies]
    bodies = [S.BodyDesc(jt[b.joint_type], np.eye(4), b.rest_E, b.inertia, b.mass) for b in sc.bodies]
    pairs = [S.PairDesc(a, b, grids[b], sc.bodies[a].points, sc.kn, sc.kt, sc.mu, sc.damping) for a, b in sc.pairs]
    sim = S.Sim(bodies, pairs, sc.h, sc.gravity)
    sim.q, sim.qdot, sim.f_ext = sc.q0[e].copy(), sc.qdot0[e].copy(), sc.f_ext[e].copy()
    return sim


@pytest.mark.parametrize("joint_type,n_envs", [("translational", 6), ("free3d-exp", 12)])
def test_oracle_rollouts_match_reference_forward(joint_type, n_envs):
    from oracle import ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    from assemble_them_all_b200 import scenes
    sc = scenes.c2_planner_rollouts(n_envs=n_envs, n_vertices=600, sdf_dx=0.1, joint_type=joint_type)
    p = ref_scene.make_probe(sc.bodies, sc.kn, sc.kt, sc.mu, sc.damping, joint_type=joint_type, inject_frames=True)
    assert p.timestep() == sc.h
    steps = 60
    touched_any = False
    for e in range(n_envs):
        _, q, qd, Q, QD = ref_scene.rollout(p, sc.q0[e], sc.qdot0[e], sc.f_ext[e, 0], steps, trajectory=True)
        sim = _oracle_for(sc, e)
        for s in range(steps):
            sim.step()
            assert_state_close(Q[s], QD[s], sim.q, sim.qdot, sc.h, f"env {e} step {s}")
        assert sim.converged == p.is_converged()
        assert bool(sim.touched.any()) == (len(p.get_contact_bodies("part0")) > 0)
        touched_any |= bool(sim.touched.any())
    assert touched_any
