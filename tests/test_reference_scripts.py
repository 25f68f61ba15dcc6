"""The reference's own planner script, run UNCHANGED (oracle/make_golden_planner.py: symlink view of
/root/reference, tests/shims for trimesh / pyquaternion), pins the repo's planners:

  * tests/golden/run_joint_plan_bfs.npz holds the status and substep path `examples/run_joint_plan.py --planner bfs`
    produced on the stock CPU redmax_py for the synthetic L-channel assembly;
  * the repo's sequential restatement of the loop and the batched BFS planner, both on the reference's CPU
    simulator, must reproduce that path BIT FOR BIT (CPU, here);
  * when /root/reference is present (build container) the script itself is re-run and must still print the
    golden status (the shims and the symlink view keep working);
  * tests/test_planner_gpu.py::test_plan_matches_the_reference_script compares the CUDA stepper with the same
    golden on the GPU box, where /root/reference does not exist."""
import os

import numpy as np
import pytest

from assemble_them_all_b200 import planner as PL
from assemble_them_all_b200 import synth
from tests.drivers import seq_planner as SP
from tests.helpers import GOLDEN


def golden():
    return np.load(os.path.join(GOLDEN, "run_joint_plan_bfs.npz"))


def test_planners_reproduce_the_reference_scripts_plan():
    from oracle import ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    z = golden()
    assert str(z["status"]) == str(z["cli_status"]) == "Success"
    parts = synth.l_channel()
    make = lambda: ref_scene.probe_from_assembly(parts, "0", ["1"], "translational", sdf_dx=float(z["sdf_dx"]))  # noqa: E731
    p = make()
    hulls = PL.HullPair(np.array(p.get_body_vertices("part0")).T, np.array(p.get_body_vertices("part1", True)).T)
    st, path, acts, stages = SP.SequentialBFSPlanner(p, "part0", hulls).plan(PL.TRANS_ACTIONS, max_depth=None)
    assert st == "Success" and np.array_equal(path, z["path"])
    pool = SP.RefSimPool(make, 6, "part0", 3, ["part0", "part1"])
    r = PL.BatchedBFSPlanner(pool, 0, hulls).plan(max_depth=None)
    assert r.status == "Success" and r.actions == acts and np.array_equal(r.path, z["path"])


def test_reference_script_runs_unchanged_with_shims():
    if not os.path.isdir("/root/reference/examples"):
        pytest.skip("/root/reference is not present on this machine")
    from oracle import make_golden_planner as MG, ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    view = MG.make_view(synth.l_channel())
    status, out = MG.run_script(view, os.path.join(MG.ROOT, "oracle", "_ref"))
    assert status == str(golden()["cli_status"]) and "planning time" in out


def golden_multi():
    return np.load(os.path.join(GOLDEN, "run_multi_plan_queue.npz"))


def test_multi_plan_script_runs_unchanged_with_shims():
    """examples/run_multi_plan.py --seq-planner queue --path-planner bfs of the reference, unchanged (symlink view,
    tests/shims), on the stock CPU redmax_py: trial order, success / failure per trial and the removal sequence equal
    the golden (a path plan that cannot succeed ends by --path-max-time, so only success / failure is compared)."""
    if not os.path.isdir("/root/reference/examples"):
        pytest.skip("/root/reference is not present on this machine")
    from oracle import make_golden_planner as MG, ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    z = golden_multi()
    view = MG.make_view_multi(synth.l_channel_with_plug())
    status, sequence, trials, out = MG.run_multi_script(view, os.path.join(MG.ROOT, "oracle", "_ref"), path_max_time=15.0)
    assert status == str(z["status"]) and sequence == [str(x) for x in z["sequence"]]
    assert [m for m, _ in trials] == [str(x) for x in z["trial_ids"]] and [ok for _, ok in trials] == [bool(x) for x in z["trial_ok"]]


def test_sequence_planner_reproduces_the_multi_plan_script():
    """The repo's SequencePlanner ("queue" variant, same use of numpy's legacy RNG) over path plans on the reference's
    CPU simulator visits the parts in the script's order and ends with the script's sequence."""
    from oracle import ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    z = golden_multi()
    parts = synth.l_channel_with_plug()

    def plan_ref(move_id, still_ids, rotation, max_depth):
        sub = {k: parts[k] for k in [move_id, *still_ids]}
        p = ref_scene.probe_from_assembly(sub, move_id, list(still_ids), "translational", sdf_dx=float(z["sdf_dx"]))
        names = [f"part{i}" for i in [move_id, *still_ids]]
        still_v = np.vstack([np.array(p.get_body_vertices(n, True)).T for n in names[1:]])
        hulls = PL.HullPair(np.array(p.get_body_vertices(names[0])).T, still_v)
        return SP.SequentialBFSPlanner(p, names[0], hulls).plan(PL.TRANS_ACTIONS, max_depth=12)[0]   # depth bound instead of a clock (the last removal takes 10 moves)

    r = PL.SequencePlanner(["0", "1", "2"], plan_ref).plan_sequence("queue", False, int(z["seed"]))
    assert r.status == str(z["status"]) and r.sequence == [str(x) for x in z["sequence"]]
    assert [m for m, _, _, _ in r.log] == [str(x) for x in z["trial_ids"]]
    assert [s in PL.SUCCESS_STATUS for _, _, _, s in r.log] == [bool(x) for x in z["trial_ok"]]
