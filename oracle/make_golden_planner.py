"""Generate tests/golden/run_joint_plan_bfs.npz and run_multi_plan_queue.npz by running the reference's OWN planner
scripts, unchanged.

TEST INFRASTRUCTURE ONLY. `examples/run_joint_plan.py --planner bfs` of the reference is executed (a) as the
command-line script and (b) through its unmodified module object (to read the path it found), on the stock CPU
`redmax_py` of oracle/_ref, with tests/shims standing in for trimesh / pyquaternion (tests/shims/README.md).
Nothing of the reference is copied: the script is run from a scratch directory of SYMLINKS to /root/reference
(the reference requires assemblies to live under its `assets/` folder, which is read-only here).

    make -C oracle/ref_build -j8 && python oracle/make_golden_planner.py

Fixture: status, substep path (every q after every forward(1), run_joint_plan.py:551-554) and the Status line the
script printed, for the depth-2 L-channel assembly of assemble_them_all_b200.synth.l_channel() at sdf_dx 0.1.
tests/test_reference_scripts.py checks that the repo's planners reproduce these decisions bit for bit on the CPU
reference simulator, tests/test_planner_gpu.py that the CUDA stepper reproduces them (integers equal, path 1e-9).
"""
from __future__ import annotations

import importlib.util
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def make_view(parts, assembly_id="00000"):
    """Scratch checkout view: symlinks to the reference's examples/ utils/ assets/*.py + the synthetic assembly."""
    from assemble_them_all_b200 import synth
    view = tempfile.mkdtemp(prefix="ref_view_")
    os.symlink(os.path.join(REF, "examples"), os.path.join(view, "examples"))
    os.symlink(os.path.join(REF, "utils"), os.path.join(view, "utils"))
    os.makedirs(os.path.join(view, "assets", "joint_assembly"))
    for f in os.listdir(os.path.join(REF, "assets")):
        if f.endswith(".py"):
            os.symlink(os.path.join(REF, "assets", f), os.path.join(view, "assets", f))
    synth.write_assembly(parts, os.path.join(view, "assets", "joint_assembly", assembly_id))
    return view


def script_env(module_dir):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "shims"), module_dir, env.get("PYTHONPATH", "")])
    env["OMP_NUM_THREADS"] = "1"
    return env


def run_script(view, module_dir, extra=()):
    """The command-line script, unchanged -> (status, stdout)."""
    r = subprocess.run([sys.executable, os.path.join(view, "examples", "run_joint_plan.py"), "--planner", "bfs", "--id", "00000",
                        "--sdf-dx", "0.1", "--max-time", "100000", *extra], capture_output=True, text=True, env=script_env(module_dir),
                       cwd=view, timeout=1800)
    m = re.search(r"Status: (\w+)", r.stdout)
    if r.returncode != 0 or not m:
        raise RuntimeError("run_joint_plan.py failed:\n" + r.stdout + r.stderr)
    return m.group(1), r.stdout


def plan_in_process(view, module_dir):
    """BFSPlanner of the unmodified module object -> (status, path)."""
    for p in (os.path.join(ROOT, "tests", "shims"), module_dir):
        if p not in sys.path:
            sys.path.insert(0, p)
    spec = importlib.util.spec_from_file_location("ref_run_joint_plan", os.path.join(view, "examples", "run_joint_plan.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    asset_folder = os.path.join(view, "assets")
    planner = mod.BFSPlanner(asset_folder, os.path.join(asset_folder, "joint_assembly", "00000"), "0", ["1"], False, "sdf", 0.1,
                             1e-2, 100, 100, False)
    status, _, path = planner.plan(100000, seed=1, return_path=True)
    return status, np.array(path)


def make_view_multi(parts, assembly_id="00000"):
    """Same view with the assembly under assets/multi_assembly (run_multi_plan.py's default --dir)."""
    from assemble_them_all_b200 import synth
    view = make_view(parts, assembly_id)
    os.makedirs(os.path.join(view, "assets", "multi_assembly"))
    synth.write_assembly(parts, os.path.join(view, "assets", "multi_assembly", assembly_id))
    return view


def run_multi_script(view, module_dir, path_max_time=20.0):
    """examples/run_multi_plan.py --seq-planner queue --path-planner bfs, unchanged -> (result status, sequence,
    [(move id, succeeded)] per trial, stdout). A path plan that cannot succeed runs into --path-max-time (status
    Timeout; the BFS has no depth limit), so only success / failure per trial is machine independent."""
    r = subprocess.run([sys.executable, os.path.join(view, "examples", "run_multi_plan.py"), "--id", "00000", "--seq-planner", "queue",
                        "--path-planner", "bfs", "--sdf-dx", "0.1", "--path-max-time", str(path_max_time), "--seq-max-time", "100000"],
                       capture_output=True, text=True, env=script_env(module_dir), cwd=view, timeout=1800)
    res = re.search(r"Result: (\w+) \| Disassembled: (\d+)/(\d+)", r.stdout)
    seq = re.search(r"Sequence: \[(.*)\]", r.stdout)
    if r.returncode != 0 or not res or not seq:
        raise RuntimeError("run_multi_plan.py failed:\n" + r.stdout + r.stderr)
    trials = [(m, st in ("Success", "Start with goal")) for m, st in re.findall(r"Move id: (\w+) \| Status: ([\w ]+?) \|", r.stdout)]
    sequence = [t.strip().strip("'\"") for t in seq.group(1).split(",") if t.strip()]
    return res.group(1), sequence, trials, r.stdout


def main_multi():
    from assemble_them_all_b200 import synth
    module_dir = os.path.join(ROOT, "oracle", "_ref")
    view = make_view_multi(synth.l_channel_with_plug())
    status, sequence, trials, out = run_multi_script(view, module_dir)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "run_multi_plan_queue.npz"), status=np.array(status),
                        sequence=np.array(sequence), trial_ids=np.array([m for m, _ in trials]),
                        trial_ok=np.array([ok for _, ok in trials]), sdf_dx=np.array(0.1), seed=np.array(1))
    print(f"run_multi_plan_queue: {status}, sequence {sequence}, trials {trials}")
    shutil.rmtree(view, ignore_errors=True)


def main():
    from assemble_them_all_b200 import synth
    module_dir = os.path.join(ROOT, "oracle", "_ref")
    view = make_view(synth.l_channel())
    status_cli, out = run_script(view, module_dir)
    status, path = plan_in_process(view, module_dir)
    assert status == status_cli == "Success"
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "run_joint_plan_bfs.npz"), status=np.array(status), path=path,
                        cli_status=np.array(status_cli), sdf_dx=np.array(0.1), frame_skip=np.array(100), force_mag=np.array(100.0))
    print(f"run_joint_plan_bfs: {status}, path of {len(path)} substeps, final q {path[-1]}")
    shutil.rmtree(view, ignore_errors=True)


if __name__ == "__main__":
    main()
    main_multi()
