"""CPU: the numpy restatement of ForceGroundContact (oracle/sim_oracle.py: ground_detect,
ground_force) and the 12-dof free-fall stepping, pinned to outputs of the UNMODIFIED
reference (tests/golden/freefall_small.npz, made by oracle/make_golden.py freefall)."""
import numpy as np

from tests.helpers import assert_close, assert_state_close, load_freefall_golden, oracle_sim_from_scene


def test_ground_detection_and_force_match_reference():
    from oracle import sim_oracle as S
    z, sc = load_freefall_golden()
    branches = set()
    for c in range(int(z["n_cases"])):
        kn, kt, mu, damping = z[f"c{c}_params"]
        for gi, g in enumerate(sc["grounds"]):
            pts = sc["bodies"][g["body"]]["points"]
            gd = S.GroundDesc(g["body"], pts, g["E_g"], kn, kt, mu, damping)
            E, phi = z[f"c{c}_g{gi}_E"], z[f"c{c}_g{gi}_phi"]
            d, act = S.ground_detect(g["E_g"], E, pts)
            ids = np.nonzero(act)[0]
            assert np.array_equal(ids, z[f"c{c}_g{gi}_active"])          # integer outputs: bit-exact
            assert np.array_equal(d[ids], z[f"c{c}_g{gi}_d"])
            fm, Km, Dm, n, n_static = S.ground_force(gd, E, phi, True)
            assert n == len(ids)
            assert_close(z[f"c{c}_g{gi}_fm"], fm, f"case {c} ground {gi} fm")
            assert_close(z[f"c{c}_g{gi}_Km"], Km, f"case {c} ground {gi} Km")
            assert_close(z[f"c{c}_g{gi}_Dm"], Dm, f"case {c} ground {gi} Dm")
            fv = S.ground_force(gd, E, phi, False)[0]
            assert_close(z[f"c{c}_g{gi}_fm_value"], fv, f"case {c} ground {gi} fm (value path)")
            if mu > 1e-8:
                branches.add("static" if n_static == n else ("dynamic" if n_static == 0 else "mixed"))
    assert len(branches) >= 2  # both Coulomb branches were exercised


def test_freefall_trajectory_follows_reference():
    z, sc = load_freefall_golden()
    sim = oracle_sim_from_scene(sc)
    Q, QD = z["traj_q"], z["traj_qdot"]
    sim.q, sim.qdot = Q[0].copy(), QD[0].copy()
    steps = 260  # ground impact of the block (~step 140) and of the peg are inside this window
    for s in range(steps):
        sim.step()
        assert_state_close(Q[s + 1], QD[s + 1], sim.q, sim.qdot, sc["h"], f"step {s}")
        assert sim.converged == bool(z["traj_converged"][s])
    assert sim.touched_ground[1]
