"""CPU: the numpy restatement of the reference's BDF1 step (oracle/sim_oracle.py)
against per-step q / qdot of the UNMODIFIED reference (tests/golden/traj_small.npz,
produced by oracle/make_golden.py through oracle/_ref/ref_probe).

Bars (tests/helpers.assert_state_close): q within 1e-9 relative (north_star); qdot =
(q1 - q0) / h amplifies the Newton tolerance (|g| < 1e-8, Simulation.cpp:1008) and
any perturbation of q by 1/h = 1000, so it is held to 1e-6 relative plus a
1e-9 |q| / h floor; Newton convergence flags and the contact history bit-exact."""
import numpy as np
import pytest

from tests.helpers import assert_state_close, load_traj_golden, oracle_sim, rel_err



@pytest.mark.parametrize("joint_type", ["translational", "free3d-exp"])
def test_trajectory_matches_reference(golden_contact, joint_type):
    T = load_traj_golden()
    key = joint_type.replace("-", "_")
    sim = oracle_sim(golden_contact, T, joint_type)
    sim.f_ext[0] = T[f"{key}_force"]
    Q, QD, CV = T[f"{key}_q"], T[f"{key}_qdot"], T[f"{key}_converged"]
    assert np.array_equal(sim.q, Q[0]) and np.array_equal(sim.qdot, QD[0])
    for s in range(int(T["steps"])):
        sim.step()
        assert_state_close(Q[s + 1], QD[s + 1], sim.q, sim.qdot, sim.h, f"step {s}")
        assert sim.converged == bool(CV[s]), f"step {s}: Newton convergence flag"
    # contact history: part0 touched part1 (CollisionDetection.cpp:143-146)
    assert sim.touched.any() and [str(x) for x in T[f"{key}_touched"]] == ["part1"]
    # final body pose = Simulation::get_body_E0i
    E, _ = sim.body_poses()
    assert rel_err(T[f"{key}_part0_E_final"], E[0]) <= 1e-9
    assert rel_err(T[f"{key}_part1_E_final"], E[1]) <= 1e-9


def test_failed_line_search_is_exercised():
    """The translational fixture contains a step whose Newton solve does not converge
    (strike logic of Simulation.cpp:1047-1057)."""
    T = load_traj_golden()
    assert (~T["translational_converged"]).sum() >= 1


def test_spherical_exp_small_angle_branch():
    from oracle import sim_oracle as S
    R, Rdot, Tm, Tdot, dR, dRdot, dT, dTdot = S.spherical_exp(np.zeros(3), np.array([0.1, -0.2, 0.3]))
    assert np.array_equal(R, np.eye(3)) and np.array_equal(Tm, np.eye(3))
    assert np.allclose(Rdot, S.skew(np.array([0.1, -0.2, 0.3])))
    # just past the switch the closed form agrees with the limit
    r = np.array([1e-7, -2e-7, 1.5e-7])
    R2, _, T2, *_ = S.spherical_exp(r, np.zeros(3))
    assert np.allclose(R2, np.eye(3) + S.skew(r), atol=1e-12) and np.allclose(T2, np.eye(3), atol=1e-6)


def test_reparam_wraps_large_rotation():
    from oracle import sim_oracle as S
    bd = S.BodyDesc(S.FREE3D_EXP, np.eye(4), np.eye(4), np.ones(6), 1.0)
    q = np.array([0.1, 0.2, 0.3, 0.0, 0.0, 5.5])     # |r| > 1.5 pi
    qd = np.array([0.0, 0.0, 0.0, 0.3, -0.1, 0.2])
    q2, qd2 = S.reparam(bd, q, qd)
    assert np.linalg.norm(q2[3:]) < 1.5 * S.PI
    assert np.allclose(S.exp3(q[3:]), S.exp3(q2[3:]), atol=1e-12)          # same rotation
    _, _, T0, *_ = S.spherical_exp(q[3:], qd[3:], False)
    _, _, T1, *_ = S.spherical_exp(q2[3:], qd2[3:], False)
    assert np.allclose(T0 @ qd[3:], T1 @ qd2[3:], atol=1e-12)              # same body angular velocity
