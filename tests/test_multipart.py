"""Multi-part assemblies (C4-style: one moving peg, a perforated plate and other pegs
fixed; 2 contact forces per still part, run_joint_plan.py:169-173) built by the
UNMODIFIED reference from OBJ files (its own loader, principal-axes body frames and
sdfgen grids), then read back and fed to (a) the numpy stepper oracle [CPU test] and
(b) the CUDA batch + stepper [GPU test]; rollouts must follow the reference's own
Simulation::forward. Skipped when oracle/_ref is not built."""
import numpy as np
import pytest

from tests.helpers import assert_state_close, rel_err


def _build_reference(joint_type):
    from oracle import ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    from assemble_them_all_b200 import synth
    parts = synth.peg_array(nx=2, ny=2, n_vertices=400, n_parts=5)
    p = ref_scene.probe_from_assembly(parts, "2", ["0", "1", "3", "4"], joint_type=joint_type, sdf_dx=0.15)
    return p, ref_scene.extract_scene(p)


def _actions(nr):
    acts = []
    for d in ([0, 0, 1], [1, 0, 0], [0, -1, 0], [0.6, 0.0, 0.8]):
        f = np.zeros(6)
        f[3:] = 100.0 * np.asarray(d, dtype=np.float64)
        acts.append(f)
    if nr == 6:
        f = np.zeros(6)
        f[:3] = [0.0, 300.0, 0.0]
        f[3:] = [0.0, 0.0, 60.0]
        acts.append(f)
    return acts


def _reference_rollout(p, f, steps):
    from oracle import ref_scene
    nr = p.ndof_r()
    return ref_scene.rollout(p, np.zeros(nr), np.zeros(nr), f, steps, trajectory=True, move_name="part2")


@pytest.mark.parametrize("joint_type", ["translational", "free3d-exp"])
def test_oracle_follows_reference_on_five_parts(joint_type):
    from oracle import sdfc_oracle as O, sim_oracle as S
    O.build()
    p, sc = _build_reference(joint_type)
    assert len(sc["bodies"]) == 5 and len(sc["forces"]) == 8
    jt = {"fixed": S.FIXED, "translational": S.TRANSLATIONAL, "free3d-exp": S.FREE3D_EXP}
    grids = [O.Grid.from_dict(b["grid"]) for b in sc["bodies"]]
    steps = 50
    for f in _actions(p.ndof_r()):
        _, q, qd, Q, QD = _reference_rollout(p, f, steps)
        bodies = [S.BodyDesc(jt[b["joint_type"]], b["E_pj_0"], b["E_ji"], b["inertia"], b["mass"]) for b in sc["bodies"]]
        pairs = [S.PairDesc(fo["a"], fo["b"], grids[fo["b"]], sc["bodies"][fo["a"]]["points"], fo["kn"], fo["kt"], fo["mu"],
                            fo["damping"]) for fo in sc["forces"]]
        sim = S.Sim(bodies, pairs, sc["h"], sc["gravity"])
        sim.f_ext[0] = f
        for s in range(steps):
            sim.step()
            assert_state_close(Q[s], QD[s], sim.q, sim.qdot, sc["h"], f"step {s}")
        assert sim.converged == p.is_converged()
        assert bool(sim.touched.any()) == (len(p.get_contact_bodies("part2")) > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("joint_type", ["translational", "free3d-exp"])
def test_cuda_stepper_follows_reference_on_five_parts(cuda_ctx, joint_type):
    from assemble_them_all_b200 import contact
    p, sc = _build_reference(joint_type)
    acts = _actions(p.ndof_r())
    n_envs, steps = len(acts), 50
    sdf = [contact.SdfGrid.from_dict(cuda_ctx, b["grid"]) for b in sc["bodies"]]
    assert all(s.is_fp32 for s in sdf)  # fresh sdfgen grids are fp32-valued (BodySDFObj.cpp:81-83)
    pts = [contact.ContactPoints(cuda_ctx, b["points"]) for b in sc["bodies"]]
    pairs = [contact.ForceGeneralSDFContact(cuda_ctx, pts[f["a"]], sdf[f["b"]], f["kn"], f["kt"], f["mu"], f["damping"],
                                            6 * f["a"], 6 * f["b"]) for f in sc["forces"]]
    batch = contact.ContactBatch(cuda_ctx, n_envs, len(sc["bodies"]), pairs, [f["a"] for f in sc["forces"]],
                                 [f["b"] for f in sc["forces"]])
    sim = contact.BatchedSimulation(batch, sc["bodies"], sc["gravity"], sc["h"])
    sim.reset()
    sim.set_body_external_force(0, np.stack(acts))
    Q, QD = sim.forward(steps, trajectory=True)
    conv, hist = sim.is_converged(), sim.get_contact_history()
    E = sim.get_body_E0i()
    for e, f in enumerate(acts):
        _, q, qd, Qr, QDr = _reference_rollout(p, f, steps)
        for s in range(steps):
            assert_state_close(Qr[s], QDr[s], Q[s, e], QD[s, e], sc["h"], f"env {e} step {s}")
        assert bool(conv[e]) == p.is_converged()
        assert bool(hist[e].any()) == (len(p.get_contact_bodies("part2")) > 0)
        for b, body in enumerate(sc["bodies"]):
            assert rel_err(p.body_info(body["name"])["E_0i"], E[e, b]) <= 1e-9
    # one teacher-forced evaluation of all 8 forces at the final poses: system arrays vs the reference's Robot loop
    Eb = np.stack([np.stack([np.asarray(p.body_info(b["name"])["E_0i"]) for b in sc["bodies"]])] * n_envs)
    phib = np.stack([np.stack([np.asarray(p.body_info(b["name"])["phi"]) for b in sc["bodies"]])] * n_envs)
    fm, Km, Dm, nact = batch.eval(Eb, phib, deriv=True)
    ref = [p.force_deriv_public(fi) for fi in range(p.num_forces())]
    from tests.helpers import assert_close
    if any(np.abs(r[0]).max() > 0 for r in ref):
        assert_close(sum(r[0] for r in ref), fm[0], "fm system (8 forces)")
        assert_close(sum(r[1] for r in ref), Km[0], "Km system (8 forces)")
    sim.close(); batch.close()
