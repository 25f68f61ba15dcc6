"""GPU: the drop-in boundary end to end. The reference's OWN host (XML parser, OBJ
loader, sdfgen, Robot, BDF1/Newton, pybind surface — unmodified objects from
/root/reference) is run twice on the same synthetic assembly, driven the way the BFS
planner drives it (tests/drivers/redmax_rollout.py):

  oracle/_ref/redmax_py        stock CPU ForceGeneralSDFContact
  oracle/_ref/b200/redmax_py   ForceGeneralSDFContactB200 -> C-ABI -> CUDA kernels

and the per-substep joint states must agree: q 1e-9 relative, qdot to the derived
bar of tests/helpers.assert_state_close, Newton convergence flags and contact
history exactly. Both modules are built by oracle/ref_build/Makefile in the build
container and travel to the GPU box; the test is skipped if they are absent."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import ROOT, assert_state_close, rel_err

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref")
DRIVER = os.path.join(ROOT, "tests", "drivers", "redmax_rollout.py")


def _run(module_dir, out, joint, nverts=600, dx=0.1, steps=100, scene="planner"):
    r = subprocess.run([sys.executable, DRIVER, module_dir, out, "--joint", joint, "--nverts", str(nverts), "--dx", str(dx),
                        "--steps", str(steps), "--scene", scene], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.load(out), r.stdout.strip()


@pytest.mark.parametrize("joint", ["translational", "free3d-exp"])
def test_reference_simulation_with_b200_force_matches_stock(tmp_path, joint):
    if not (glob.glob(os.path.join(REF, "redmax_py*.so")) and glob.glob(os.path.join(REF, "b200", "redmax_py*.so"))):
        pytest.skip("oracle/_ref (stock + b200 redmax_py) not built")
    stock, msg_s = _run(REF, str(tmp_path / "stock.npz"), joint)
    b200, msg_b = _run(os.path.join(REF, "b200"), str(tmp_path / "b200.npz"), joint)
    print(msg_s)
    print(msg_b)
    h = float(np.float32(1e-3))
    n_actions, steps = stock["Q"].shape[:2]
    for a in range(n_actions):
        for s in range(steps):
            assert_state_close(stock["Q"][a, s], stock["QD"][a, s], b200["Q"][a, s], b200["QD"][a, s], h,
                               f"action {a} step {s}")
    assert np.array_equal(stock["converged"], b200["converged"])
    assert [str(t) for t in stock["touched"]] == [str(t) for t in b200["touched"]]
    assert any(str(t) == "part1" for t in stock["touched"])
    assert rel_err(stock["E_final"], b200["E_final"]) <= 1e-9


def test_reference_freefall_with_b200_ground_sdf_contact_and_sdfgen_matches_stock(tmp_path):
    """test_joint_sim.py's configuration driven through the reference's own Simulation: in the b200
    module ForceGroundContactB200, ForceGeneralSDFContactB200 and the GPU sdfgen (BodySDFObj and
    SDFMesh call sites) replace the CPU code. Grids are bit-identical, so SDFMesh.distance is
    compared bit for bit; the 300-step trajectory to 1e-9."""
    if not (glob.glob(os.path.join(REF, "redmax_py*.so")) and glob.glob(os.path.join(REF, "b200", "redmax_py*.so"))):
        pytest.skip("oracle/_ref (stock + b200 redmax_py) not built")
    stock, msg_s = _run(REF, str(tmp_path / "stock.npz"), "free3d-exp", steps=300, scene="freefall")
    b200, msg_b = _run(os.path.join(REF, "b200"), str(tmp_path / "b200.npz"), "free3d-exp", steps=300, scene="freefall")
    print(msg_s)
    print(msg_b)
    h = float(np.float32(1e-3))

    # the reference's CPU path is deterministic: any stock != b200 difference is a failure of this repo
    for s in range(stock["Q"].shape[1]):
        assert_state_close(stock["Q"][0, s], stock["QD"][0, s], b200["Q"][0, s], b200["QD"][0, s], h, f"step {s}")
    assert np.array_equal(stock["sdfmesh_distance"], b200["sdfmesh_distance"])   # GPU sdfgen grid + GPU sampling (sdfc_sdf_distance)
    assert float(stock["sdfmesh_min_distance"]) == float(b200["sdfmesh_min_distance"])
    assert (stock["sdfmesh_distance"] < 1e30).any() and (stock["sdfmesh_distance"] > 1e30).any()  # inside and outside the box
    assert np.array_equal(stock["converged"], b200["converged"])
    assert [str(t) for t in stock["touched"]] == [str(t) for t in b200["touched"]]
    assert "ground" in str(stock["touched"][1])
    assert rel_err(stock["E_final"], b200["E_final"]) <= 1e-9


def test_b200_module_fails_loudly_without_the_library(tmp_path):
    """No CPU fallback behind the plug-in: with the CUDA device hidden the Simulation
    constructor (Robot::init -> Force::init) throws instead of falling back."""
    if not glob.glob(os.path.join(REF, "b200", "redmax_py*.so")):
        pytest.skip("oracle/_ref/b200 not built")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, DRIVER, os.path.join(REF, "b200"), str(tmp_path / "x.npz")], capture_output=True,
                       text=True, timeout=600, env=env)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)
