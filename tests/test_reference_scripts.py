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
