"""GPU: planner decisions. The batched BFS planner on the CUDA stepper (through the C-ABI)
against the sequential planner loop on the UNMODIFIED reference simulator
(oracle/_ref/ref_probe, CPU) on the same synthetic assemblies, built by the reference
itself from OBJ files. Integer outcomes (status, action sequence, stages, rollouts
consumed, substeps per rollout) must be identical; the substep path agrees to 1e-9
relative for the translational plans. Free-running 6-dof rollouts that grind along a wall
for hundreds of steps amplify last-bit differences of the force sums (kn = 1e6), so the
rotational rollouts use a stated free-running drift bar of 1e-6 (the per-step 1e-9 bar is
held by tests/test_sim_gpu.py with the reference's own trajectories)."""
FREE_RUNNING_6DOF_RTOL = 1e-6
import numpy as np
import pytest

from assemble_them_all_b200 import planner as PL
from assemble_them_all_b200 import synth
from tests.drivers import seq_planner as SP
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _setup(cuda_ctx, parts, joint_type, dx, n_envs):
    from assemble_them_all_b200 import contact
    from oracle import ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    p = ref_scene.probe_from_assembly(parts, "0", ["1"], joint_type, sdf_dx=dx)
    hulls = PL.HullPair(np.array(p.get_body_vertices("part0")).T, np.array(p.get_body_vertices("part1", True)).T)
    sim = contact.build_simulation(cuda_ctx, ref_scene.extract_scene(p), n_envs)
    return p, hulls, sim


@pytest.mark.parametrize("n_envs", [6, 64])
def test_translational_depth2_plan_matches_reference(cuda_ctx, n_envs):
    p, hulls, sim = _setup(cuda_ctx, synth.l_channel(), "translational", 0.1, n_envs)
    seq = SP.SequentialBFSPlanner(p, "part0", hulls)
    st, path, acts, stages = seq.plan(PL.TRANS_ACTIONS, max_depth=8)
    assert st == "Success" and acts == [4, 0]
    r = PL.BatchedBFSPlanner(sim, 0, hulls).plan(max_depth=8)
    assert (r.status, r.actions, r.stages, r.rollouts) == (st, acts, stages, seq.rollouts)
    assert r.path.shape == path.shape
    assert rel_err(path, r.path) <= 1e-9
    if n_envs == 64:  # the whole second level ran as one wave
        assert r.forward_calls < 30
    sim.close()


def test_action_order_on_reoriented_peg(cuda_ctx):
    R = synth.axis_angle_rotation(np.array([0.0, 1.0, 0.0]), -np.pi / 2)
    parts = synth.reorient(synth.peg_in_hole(n_vertices=1500, peg_z=(-1.0, 4.0)), R)
    p, hulls, sim = _setup(cuda_ctx, parts, "translational", 0.1, 6)
    seq = SP.SequentialBFSPlanner(p, "part0", hulls)
    st, path, acts, stages = seq.plan(PL.TRANS_ACTIONS, max_depth=1)
    r = PL.BatchedBFSPlanner(sim, 0, hulls).plan(max_depth=1)
    assert (r.status, r.actions, r.stages, r.rollouts) == (st, acts, stages, seq.rollouts)
    assert st == "Success" and r.path.shape == path.shape and rel_err(path, r.path) <= 1e-9
    sim.close()


def test_rotational_level_matches_reference(cuda_ctx):
    """12 actions (6 translations, 6 torques) from the start state; every rollout's stop
    chunk and final state must match the reference rollout by rollout."""
    p, hulls, sim = _setup(cuda_ctx, synth.l_channel(), "free3d-exp", 0.1, 12)
    seq = SP.SequentialBFSPlanner(p, "part0", hulls, rotation=True)
    ref = [seq._rollout(np.zeros(6), [], a) for a in PL.ROT_ACTIONS]
    bp = PL.BatchedBFSPlanner(sim, 0, hulls, rotation=True)
    got = bp._run_wave([(np.zeros(6), np.zeros((0, 6)), a) for a in range(12)])
    for a in range(12):
        st, new_q, temp_path = ref[a]
        gst, gpath, chunks = got[a]
        assert gst == st and len(temp_path) == len(gpath) == 100 * chunks, f"action {a}"
        assert rel_err(np.array(temp_path), gpath) <= FREE_RUNNING_6DOF_RTOL, f"action {a}"
    r = bp.plan(max_depth=1)
    assert r.status == "Failure" and r.rollouts == 12
    sim.close()


def test_inactive_environments_keep_their_state(cuda_ctx):
    p, hulls, sim = _setup(cuda_ctx, synth.l_channel(), "translational", 0.1, 4)
    q0 = np.array([[0.0, 0, 0], [0.1, 0, 0], [0.2, 0, 0], [0.3, 0, 0]])
    sim.reset(q0, np.zeros_like(q0))
    sim.set_body_external_force(0, np.array([0, 0, 0, 100.0, 0, 0]))
    sim.set_active([True, False, True, False])
    sim.forward(10)
    q, qd = sim.get_state()
    assert np.array_equal(q[[1, 3]], q0[[1, 3]]) and not qd[[1, 3]].any()
    assert (q[[0, 2], 0] > q0[[0, 2], 0]).all()
    assert sim.stats()["env_steps"] == 20
    sim.set_active(None)
    sim.forward(1)
    assert sim.stats()["env_steps"] == 24
    sim.close()


def test_sequence_plan_matches_reference(cuda_ctx):
    """C4-style sequence planning (ProgressiveQueueSequencePlanner, run_multi_plan.py:228-330) on a
    three-part assembly: every candidate removal is one batched BFS plan on the CUDA stepper; trial
    order, per-trial status and the final removal sequence must equal the reference loop's."""
    from assemble_them_all_b200 import contact
    from oracle import ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    parts = synth.reorient(synth.l_channel_with_plug(), np.array([[0.0, 1, 0], [0, 0, 1], [1, 0, 0]]))
    cache = {}

    def scene(move_id, still_ids):
        key = (move_id, tuple(still_ids))
        if key not in cache:
            sub = {k: parts[k] for k in [move_id, *still_ids]}
            p = ref_scene.probe_from_assembly(sub, move_id, list(still_ids), "translational", sdf_dx=0.1)
            names = [f"part{i}" for i in [move_id, *still_ids]]
            still_v = np.vstack([np.array(p.get_body_vertices(n, True)).T for n in names[1:]])
            hulls = PL.HullPair(np.array(p.get_body_vertices(names[0])).T, still_v)
            cache[key] = (p, hulls, names)
        return cache[key]

    def plan_ref(move_id, still_ids, rotation, max_depth):
        p, hulls, names = scene(move_id, still_ids)
        return SP.SequentialBFSPlanner(p, names[0], hulls).plan(PL.TRANS_ACTIONS, max_depth=max_depth)[0]

    def plan_cuda(move_id, still_ids, rotation, max_depth):
        p, hulls, names = scene(move_id, still_ids)
        sim = contact.build_simulation(cuda_ctx, ref_scene.extract_scene(p), 12)
        try:
            return PL.BatchedBFSPlanner(sim, 0, hulls).plan(max_depth=max_depth).status
        finally:
            sim.close()

    st, seq, count, trace = SP.reference_prog_queue(["0", "1", "2"], plan_ref, False, 1, max_trials=12)
    r = PL.SequencePlanner(["0", "1", "2"], plan_cuda).plan_sequence("prog-queue", False, 1, max_trials=12)
    assert (r.status, r.sequence, r.trials) == (st, seq, count)
    assert [(m, d, s) for m, d, _, s in r.log] == trace
    assert st == "Success" and seq == ["2", "0"]


def test_planner_wall_time_report(cuda_ctx):
    """One whole planning problem at a realistic mesh size (L-channel, ~13 k contact points on the
    block, dx 0.05): sequential reference loop on the CPU vs the batched planner on the GPU. Same
    plan required; wall times are printed (kept under profiles/ by the profiling script)."""
    import time
    p, hulls, sim = _setup(cuda_ctx, synth.l_channel(h=0.1), "translational", 0.05, 36)
    npts = [len(np.array(p.get_body_vertices(n)).T) for n in ("part0", "part1")]
    seq = SP.SequentialBFSPlanner(p, "part0", hulls)
    t0 = time.perf_counter()
    st, path, acts, stages = seq.plan(PL.TRANS_ACTIONS, max_depth=8)
    t_cpu = time.perf_counter() - t0
    bp = PL.BatchedBFSPlanner(sim, 0, hulls)
    bp.plan(max_depth=8)                                    # warm-up (first-use allocations)
    t0 = time.perf_counter()
    r = bp.plan(max_depth=8)
    t_gpu = time.perf_counter() - t0
    print(f"\nPLANNER: contact points {npts}; plan {acts} ({st}); reference loop: {seq.rollouts} rollouts, {seq.env_steps} env-steps, "
          f"{t_cpu:.2f} s on one core; batched planner: {r.rollouts_run} rollouts run ({r.rollouts} consumed), {r.env_steps} env-steps, "
          f"{r.forward_calls} forward calls, {t_gpu:.3f} s -> {t_cpu / t_gpu:.1f}x")
    assert (r.status, r.actions, r.stages, r.rollouts) == (st, acts, stages, seq.rollouts)
    assert r.path.shape == path.shape and rel_err(path, r.path) <= 1e-9
    sim.close()


def test_plan_matches_the_reference_script(cuda_ctx):
    """tests/golden/run_joint_plan_bfs.npz: status and substep path of the reference's UNMODIFIED
    `examples/run_joint_plan.py --planner bfs` on the stock CPU simulator (oracle/make_golden_planner.py).
    The batched planner on the CUDA stepper must take the same decisions (status, number of substeps = which
    rollout ended the search and after how many frame_skip chunks) and follow the same path to 1e-9."""
    import os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "run_joint_plan_bfs.npz"))
    p, hulls, sim = _setup(cuda_ctx, synth.l_channel(), "translational", float(z["sdf_dx"]), 64)
    r = PL.BatchedBFSPlanner(sim, 0, hulls).plan(max_depth=None)
    assert r.status == str(z["status"]) == "Success"
    assert r.path.shape == z["path"].shape
    assert rel_err(z["path"], r.path) <= 1e-9
    sim.close()


def test_sequence_matches_the_reference_multi_plan_script(cuda_ctx):
    """tests/golden/run_multi_plan_queue.npz is what the reference's own examples/run_multi_plan.py (queue + bfs,
    unchanged, stock CPU redmax_py) printed for the three-part assembly (oracle/make_golden_planner.py). The
    repo's SequencePlanner over batched BFS plans on the CUDA stepper must visit the parts in the same order,
    succeed / fail on the same trials and end with the same removal sequence."""
    import os
    from assemble_them_all_b200 import contact
    from oracle import ref_scene
    from tests.helpers import GOLDEN
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    z = np.load(os.path.join(GOLDEN, "run_multi_plan_queue.npz"))
    parts = synth.l_channel_with_plug()

    def plan_cuda(move_id, still_ids, rotation, max_depth):
        sub = {k: parts[k] for k in [move_id, *still_ids]}
        p = ref_scene.probe_from_assembly(sub, move_id, list(still_ids), "translational", sdf_dx=float(z["sdf_dx"]))
        names = [f"part{i}" for i in [move_id, *still_ids]]
        still_v = np.vstack([np.array(p.get_body_vertices(n, True)).T for n in names[1:]])
        hulls = PL.HullPair(np.array(p.get_body_vertices(names[0])).T, still_v)
        sim = contact.build_simulation(cuda_ctx, ref_scene.extract_scene(p), 12)
        try:
            return PL.BatchedBFSPlanner(sim, 0, hulls).plan(max_depth=12).status   # depth bound instead of a clock
        finally:
            sim.close()

    r = PL.SequencePlanner(["0", "1", "2"], plan_cuda).plan_sequence("queue", False, int(z["seed"]))
    assert r.status == str(z["status"]) and r.sequence == [str(x) for x in z["sequence"]]
    assert [m for m, _, _, _ in r.log] == [str(x) for x in z["trial_ids"]]
    assert [s in PL.SUCCESS_STATUS for _, _, _, s in r.log] == [bool(x) for x in z["trial_ok"]]
