"""Generate tests/golden/sdf_cache_small.{sdf,npz} from the UNMODIFIED reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY. Pins the `.sdf` text cache (the only on-disk format next to the hot
path) to the reference's own writer and reader:

  1. a reference Simulation is constructed with `load_sdf="true" save_sdf="true"` on a small
     synthetic pair, so BodySDFObj::save_SDF (Body/BodySDFObj.cpp:105-116) writes `<obj>.sdf`;
  2. a second construction takes the BodySDFObj::load_SDF branch (Body/BodySDFObj.cpp:86-103)
     and holds the reloaded (fp64, 6 significant digits) grid.

Committed: the file the reference wrote (bytes), the grid it was written from (fp32-valued) and
the grid / box / spacing the reference read back. Run from the repo root in the build container:

    make -C oracle/ref_build -j8 && python oracle/make_golden_sdf_cache.py
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from assemble_them_all_b200 import synth  # noqa: E402

import ref_probe  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    tmp = tempfile.mkdtemp(prefix="sdf_cache_")
    parts = synth.peg_in_hole(n_vertices=300)
    synth.write_assembly(parts, tmp)
    xml = synth.planner_xml(tmp, "0", ["1"], sdf_dx=0.4, save_sdf=True)
    p1 = ref_probe.Probe(xml, tmp)                 # precompute_SDF + save_SDF
    assert os.path.exists(os.path.join(tmp, "0.sdf")) and os.path.exists(os.path.join(tmp, "1.sdf"))
    p2 = ref_probe.Probe(xml, tmp)                 # load_SDF
    out = {}
    name, pid = "part0", "0"
    g1, g2 = dict(p1.sdf_grid(name)), dict(p2.sdf_grid(name))
    for tag, g in (("fresh", g1), ("reloaded", g2)):
        out[f"{tag}_dims"] = np.array([g["ni"], g["nj"], g["nk"]], dtype=np.int32)
        out[f"{tag}_min_box"] = np.asarray(g["min_box"], dtype=np.float64)
        out[f"{tag}_max_box"] = np.asarray(g["max_box"], dtype=np.float64)
        out[f"{tag}_spacing"] = np.array([g["dx"], g["dy"], g["dz"]], dtype=np.float64)
        out[f"{tag}_grid"] = np.asarray(g["grid"], dtype=np.float64).reshape(-1)
    assert np.array_equal(out["fresh_grid"], out["fresh_grid"].astype(np.float32).astype(np.float64))
    assert not np.array_equal(out["fresh_grid"], out["reloaded_grid"])
    os.makedirs(GOLDEN, exist_ok=True)
    shutil.copyfile(os.path.join(tmp, f"{pid}.sdf"), os.path.join(GOLDEN, "sdf_cache_small.sdf"))
    np.savez_compressed(os.path.join(GOLDEN, "sdf_cache_small.npz"), **out)
    n = int(np.prod(out["fresh_dims"]))
    print(f"sdf_cache_small: {name} grid {tuple(out['fresh_dims'])} = {n} values, "
          f"{os.path.getsize(os.path.join(GOLDEN, 'sdf_cache_small.sdf'))} bytes of text")


if __name__ == "__main__":
    main()
