"""GPU parity of the batched BDF1 stepper (sdfc_sim_*, through the C-ABI) against
the reference's own trajectories (tests/golden/traj_small.npz) and the CPU oracle.
Bars (tests/helpers.assert_state_close): q 1e-9 relative; qdot 1e-6 relative plus
the 1e-9 |q| / h floor of a derived quantity; convergence flags / contact history exact."""
import numpy as np
import pytest

from tests.helpers import assert_state_close, load_traj_golden, oracle_sim, rel_err, traj_bodies

pytestmark = pytest.mark.gpu


def build(ctx, G, T, joint_type, n_envs, kn=1e6, kt=1e3, mu=0.0, damping=0.0):
    from assemble_them_all_b200 import contact
    sdf = {n: contact.SdfGrid.from_dict(ctx, G.grid_dict(n)) for n in G.names}
    pts = {n: contact.ContactPoints(ctx, G.points(n)) for n in G.names}
    pairs = [contact.ForceGeneralSDFContact(ctx, pts["part0"], sdf["part1"], kn, kt, mu, damping, 0, 6),
             contact.ForceGeneralSDFContact(ctx, pts["part1"], sdf["part0"], kn, kt, mu, damping, 6, 0)]
    batch = contact.ContactBatch(ctx, n_envs, 2, pairs, [0, 1], [1, 0])
    key = joint_type.replace("-", "_")
    sim = contact.BatchedSimulation(batch, traj_bodies(T, joint_type), [0.0, 0.0, 1e-12], float(T[f"{key}_h"]))
    return sim


@pytest.mark.parametrize("joint_type", ["translational", "free3d-exp"])
def test_trajectory_matches_reference_golden(cuda_ctx, golden_contact, joint_type):
    T = load_traj_golden()
    key = joint_type.replace("-", "_")
    steps = int(T["steps"])
    sim = build(cuda_ctx, golden_contact, T, joint_type, n_envs=3)
    assert sim.ndof_r == (3 if joint_type == "translational" else 6)
    sim.reset()
    sim.set_body_external_force(0, T[f"{key}_force"])
    Q, QD = sim.forward(steps, trajectory=True)
    Qr, QDr, CV = T[f"{key}_q"], T[f"{key}_qdot"], T[f"{key}_converged"]
    for s in range(steps):
        for e in range(3):
            assert_state_close(Qr[s + 1], QDr[s + 1], Q[s, e], QD[s, e], sim.h, f"step {s} env {e}")
    # identical environments are bit-identical (fixed-order reductions)
    assert np.array_equal(Q[:, 0], Q[:, 1]) and np.array_equal(QD[:, 0], QD[:, 2])
    q, qd = sim.get_state()
    assert np.array_equal(q, Q[-1]) and np.array_equal(qd, QD[-1])
    assert np.array_equal(sim.is_converged(), np.full(3, bool(CV[-1])))
    hist = sim.get_contact_history()
    assert hist[:, 0].all()
    E = sim.get_body_E0i()
    assert rel_err(T[f"{key}_part0_E_final"], E[0, 0]) <= 1e-9
    assert rel_err(T[f"{key}_part1_E_final"], E[0, 1]) <= 1e-9
    st = sim.stats()
    assert st["env_steps"] == 3 * steps and st["deriv_env_evals"] >= 3 * steps
    sim.close()


def test_forward_in_chunks_equals_one_call(cuda_ctx, golden_contact):
    """forward(100) == 10 x forward(10): the planner's frame_skip loop (run_joint_plan.py:551-554)."""
    T = load_traj_golden()
    sim = build(cuda_ctx, golden_contact, T, "translational", n_envs=2)
    sim.reset()
    sim.set_body_external_force(0, T["translational_force"])
    Q1, _ = sim.forward(60, trajectory=True)
    sim.reset()
    sim.set_body_external_force(0, T["translational_force"])
    chunks = [sim.forward(10, trajectory=True)[0] for _ in range(6)]
    assert np.array_equal(Q1, np.concatenate(chunks))
    sim.close()


@pytest.mark.parametrize("joint_type,mu,damping", [("translational", 0.0, 0.0), ("free3d-exp", 0.0, 0.0),
                                                  ("free3d-exp", 0.3, 10.0)])
def test_heterogeneous_envs_match_oracle(cuda_ctx, golden_contact, joint_type, mu, damping):
    """Every environment gets its own push and start state, so Newton iteration and
    line-search counts differ across the batch; each must follow the CPU oracle."""
    T = load_traj_golden()
    rng = np.random.default_rng(7)
    n_envs, steps = 6, 40
    sim = build(cuda_ctx, golden_contact, T, joint_type, n_envs, mu=mu, damping=damping)
    nr = sim.ndof_r
    q0 = rng.uniform(-0.02, 0.02, size=(n_envs, nr))
    qd0 = rng.uniform(-0.3, 0.3, size=(n_envs, nr))
    if nr == 6:
        q0[:, 3:] *= 0.5
        q0[0, 3:] = 0.0  # small-angle branch of JointSphericalExp::update
    force = np.zeros((n_envs, 6))
    force[:, 3:] = rng.uniform(-120, 120, size=(n_envs, 3))
    if nr == 6:
        force[:, :3] = rng.uniform(-60, 60, size=(n_envs, 3))
    sim.reset(q0, qd0)
    sim.set_body_external_force(0, force)
    Q, QD = sim.forward(steps, trajectory=True)
    conv = sim.is_converged()
    hist = sim.get_contact_history()
    for e in range(n_envs):
        ref = oracle_sim(golden_contact, T, joint_type, mu=mu, damping=damping)
        ref.q, ref.qdot = q0[e].copy(), qd0[e].copy()
        ref.f_ext[0] = force[e]
        for s in range(steps):
            ref.step()
            assert_state_close(ref.q, ref.qdot, Q[s, e], QD[s, e], sim.h, f"env {e} step {s}")
        assert bool(conv[e]) == ref.converged
        assert np.array_equal(hist[e], ref.touched)
    sim.close()


def test_large_rotation_reparam(cuda_ctx, golden_contact):
    """|r| beyond 1.5 pi is wrapped at the end of the step (JointSphericalExp.cpp:237-264)."""
    T = load_traj_golden()
    sim = build(cuda_ctx, golden_contact, T, "free3d-exp", n_envs=2)
    q0 = np.zeros((2, 6)); qd0 = np.zeros((2, 6))
    q0[:, 2] = 30.0                      # far away from part1: free flight
    q0[0, 3:] = [0.0, 0.0, 4.70]         # just below 1.5 pi = 4.712
    q0[1, 3:] = [0.3, 0.0, 1.0]
    qd0[:, 3:] = [0.0, 0.0, 40.0]        # crosses the threshold within a few steps
    sim.reset(q0, qd0)
    Q, QD = sim.forward(5, trajectory=True)
    ref = oracle_sim(golden_contact, T, "free3d-exp")
    ref.q, ref.qdot = q0[0].copy(), qd0[0].copy()
    wrapped = False
    for s in range(5):
        ref.step()
        wrapped |= np.linalg.norm(ref.q[3:]) < 4.0
        assert_state_close(ref.q, ref.qdot, Q[s, 0], QD[s, 0], sim.h, f"step {s}")
    assert wrapped
    sim.close()


def test_rejects_unsupported_models(cuda_ctx, golden_contact):
    from assemble_them_all_b200 import contact
    T = load_traj_golden()
    bodies = traj_bodies(T, "translational")
    bodies[1]["joint_type"] = 7
    sim = None
    with pytest.raises(contact.SdfContactError):
        G = golden_contact
        sdf = contact.SdfGrid.from_dict(cuda_ctx, G.grid_dict("part1"))
        pts = contact.ContactPoints(cuda_ctx, G.points("part0"))
        pair = contact.ForceGeneralSDFContact(cuda_ctx, pts, sdf, 1e6, 1e3)
        batch = contact.ContactBatch(cuda_ctx, 1, 2, [pair], [0], [1])
        sim = contact.BatchedSimulation(batch, bodies, [0, 0, 0], 1e-3)
    assert sim is None


@pytest.mark.parametrize("joint_type,n_envs", [("translational", 12), ("free3d-exp", 12)])
def test_rollouts_match_reference_forward(cuda_ctx, joint_type, n_envs):
    """The planner rollouts bench.py times (small mesh): batched stepper against the
    UNMODIFIED reference's Simulation::forward with the same injected inputs."""
    from oracle import ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    from assemble_them_all_b200 import scenes
    sc = scenes.c2_planner_rollouts(n_envs=n_envs, n_vertices=600, sdf_dx=0.1, joint_type=joint_type)
    sim, batch, pairs, sdfs, pts = sc.upload(cuda_ctx)
    sim.reset(sc.q0, sc.qdot0)
    for b in range(len(sc.bodies)):
        sim.set_body_external_force(b, sc.f_ext[:, b])
    steps = 100
    Q, QD = sim.forward(steps, trajectory=True)
    conv = sim.is_converged()
    hist = sim.get_contact_history()
    p = ref_scene.make_probe(sc.bodies, sc.kn, sc.kt, sc.mu, sc.damping, joint_type=joint_type, inject_frames=True)
    for e in range(n_envs):
        _, q, qd, Qr, QDr = ref_scene.rollout(p, sc.q0[e], sc.qdot0[e], sc.f_ext[e, 0], steps, trajectory=True)
        for s in range(steps):
            assert_state_close(Qr[s], QDr[s], Q[s, e], QD[s, e], sim.h, f"env {e} step {s}")
        assert bool(conv[e]) == p.is_converged()
        assert bool(hist[e].any()) == (len(p.get_contact_bodies("part0")) > 0)
    sim.close()


def test_lanes_and_mixed_mode_rounds_do_not_change_results(cuda_ctx):
    """768 rollouts with different pushes drift apart in Newton / line-search phase, so most
    rounds mix value-only and derivative environments. Stepping them on 1, 2 or 3 lanes
    (environment ranges on their own streams, SDFC_SIM_LANES) must give bit-identical states:
    guards the per-(env, pair, split) slot layout of the force partials (value and derivative
    variants share the buffer inside one round) and the lane fork / join ordering."""
    import os
    from assemble_them_all_b200 import contact, scenes
    rs = scenes.c2_planner_rollouts(n_envs=768, n_vertices=600, sdf_dx=0.1)
    rs.q0[:, 2] = np.linspace(0.2, 1.5, 768)
    outs, evals = [], []
    old = os.environ.get("SDFC_SIM_LANES")
    try:
        for lanes in ("1", "2", "3"):
            os.environ["SDFC_SIM_LANES"] = lanes
            sim, *keep = rs.upload(cuda_ctx)
            sim.reset(rs.q0, rs.qdot0)
            for b in range(len(rs.bodies)):
                sim.set_body_external_force(b, rs.f_ext[:, b])
            s0 = sim.stats()
            sim.forward(60)
            s1 = sim.stats()
            outs.append(sim.get_state())
            evals.append((s1["deriv_env_evals"] - s0["deriv_env_evals"], s1["value_env_evals"] - s0["value_env_evals"]))
            assert sim.get_contact_history().any()
            sim.close()
    finally:
        if old is None:
            os.environ.pop("SDFC_SIM_LANES", None)
        else:
            os.environ["SDFC_SIM_LANES"] = old
    assert evals[0] == evals[1] == evals[2]
    assert evals[0][0] > 768 * 60          # more than one Newton iteration per step somewhere: phases diverged
    for q, qd in outs[1:]:
        assert np.array_equal(q, outs[0][0]) and np.array_equal(qd, outs[0][1])


@pytest.mark.parametrize("joint_type,frame_skip", [("translational", 100), ("free3d-exp", 80)])
def test_skin_culling_does_not_change_results(cuda_ctx, joint_type, frame_skip):
    """SDFC_SKIN=1 lets the detection kernel skip a 32-point word while Lipschitz x (displacement of the
    word since the margins were measured) stays below the word's smallest d (BatchDev::skin_*). The cull
    is conservative, so masks, contact counts and therefore every state must be BIT-identical to the
    full evaluation (SDFC_SKIN=0), for rollouts that slide freely, hit the wall and stay in contact."""
    import os
    from assemble_them_all_b200 import scenes
    rs = scenes.c2_planner_rollouts(n_envs=192, n_vertices=1500, sdf_dx=0.1, joint_type=joint_type)
    rs.q0[:, 2] = np.linspace(0.0, 2.5, 192)
    old = os.environ.get("SDFC_SKIN")
    outs = []
    try:
        for skin in ("0", "1"):
            os.environ["SDFC_SKIN"] = skin
            sim, batch, *keep = rs.upload(cuda_ctx)
            sim.reset(rs.q0, rs.qdot0)
            for b in range(len(rs.bodies)):
                sim.set_body_external_force(b, rs.f_ext[:, b])
            Q, QD = sim.forward(frame_skip, trajectory=True)
            st = sim.stats()
            hist = sim.get_contact_history()
            # a second forward from a reset: margins measured at the END of the first one are still valid state
            sim.reset(rs.q0, rs.qdot0)
            Q2, _ = sim.forward(20, trajectory=True)
            outs.append((Q, QD, Q2, hist, sim.is_converged(),
                         (st["deriv_env_evals"], st["value_env_evals"], st["rounds"]), batch.skin_stats()))
            sim.close()
    finally:
        if old is None:
            os.environ.pop("SDFC_SKIN", None)
        else:
            os.environ["SDFC_SKIN"] = old
    a, b = outs
    assert a[3].any() and not a[3].all()
    assert a[5] == b[5]
    assert not any(a[6].values())
    print("\nskin:", b[6], "of", (a[5][0] + a[5][1]) * sum(-(-len(rs.bodies[p].points) // 32) for p, _ in rs.pairs), "word visits")
    assert b[6]["words_culled"] > 0 and b[6]["rebuilds"] >= 192 * 2 and b[6]["pairs_skipped"] > 0
    for x, y in zip(a[:5], b[:5]):
        assert np.array_equal(x, y)
