"""SDF construction: the C restatement of sdfgen::make_level_set3
(oracle/sdfgen_oracle.c) and the CUDA drop-in (sdfc_make_level_set3) against grids produced
by the UNMODIFIED reference (tests/golden/sdfgen_small.npz: BodySDFObj::precompute_SDF for
exact_band 1, sdfgen::make_level_set3 called directly for exact_band 2). Every comparison
is bit for bit: the grid feeds the contact query, whose cell values must not move."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN


def _golden():
    return np.load(os.path.join(GOLDEN, "sdfgen_small.npz"))


def _args(z, n):
    sp = z[f"{n}_spacing"].astype(np.float32)
    ni, nj, nk = (int(v) for v in z[f"{n}_dims"])
    return z[f"{n}_F"], z[f"{n}_V"].astype(np.float32), z[f"{n}_min_box"].astype(np.float32), sp[0], sp[1], sp[2], ni, nj, nk


def test_oracle_level_set_matches_reference_grids():
    from oracle import sdfgen_oracle as SG
    z = _golden()
    for n in (str(x) for x in z["names"]):
        assert np.array_equal(SG.make_level_set3(*_args(z, n), exact_band=1), z[f"{n}_grid_f32"]), n
        assert np.array_equal(SG.make_level_set3(*_args(z, n), exact_band=2), z[f"{n}_grid_band2_f32"]), n
        assert (z[f"{n}_grid_f32"] < 0).any() and (z[f"{n}_grid_f32"] > 0).any()


def test_grid_layout_matches_reference():
    """Spacing / padded fp32 box / dims of BodySDFObj.cpp:41-70: oracle and the library's host entry."""
    from assemble_them_all_b200 import contact
    from oracle import sdfgen_oracle as SG
    z = _golden()
    dx = float(z["dx_request"])
    for n in (str(x) for x in z["names"]):
        for lay in (SG.grid_layout(z[f"{n}_V"], dx), contact.sdf_layout(z[f"{n}_V"], dx)):
            assert np.array_equal(lay["dims"], z[f"{n}_dims"])
            assert np.array_equal(lay["spacing"], z[f"{n}_spacing"])
            assert np.array_equal(lay["min_box"].astype(np.float64), z[f"{n}_min_box"])
            assert np.array_equal(lay["max_box"].astype(np.float64), z[f"{n}_max_box"])


def test_sdf_text_cache_round_trip(tmp_path):
    """save_SDF writes 6 significant digits (BodySDFObj.cpp:103-116): the reloaded grid equals the
    values the reference's own `infile >> double` would read back from the same text."""
    from assemble_them_all_b200 import contact
    z = _golden()
    n = str(z["names"][1])
    ni, nj, nk = (int(v) for v in z[f"{n}_dims"])
    sp = z[f"{n}_spacing"]
    g = {"ni": ni, "nj": nj, "nk": nk, "min_box": z[f"{n}_min_box"], "max_box": z[f"{n}_max_box"], "dx": sp[0], "dy": sp[1],
         "dz": sp[2], "col_th": 0.0, "grid": z[f"{n}_grid_f32"].astype(np.float64)}
    path = tmp_path / "part.sdf"
    contact.save_sdf_text(path, g)
    tokens = open(path).read().split()
    assert tokens[:3] == [str(ni), str(nj), str(nk)] and len(tokens) == 12 + ni * nj * nk
    assert tokens[12] == "%g" % g["grid"][0]  # operator<<(double): %g with precision 6
    back = contact.load_sdf_text(path, col_th=0.01)
    assert (back["ni"], back["nj"], back["nk"]) == (ni, nj, nk) and back["col_th"] == 0.01
    expect = np.array([float(t) for t in tokens[12:]])
    assert np.array_equal(back["grid"], expect)
    assert np.allclose(back["grid"], g["grid"], rtol=1e-5, atol=0) and not np.array_equal(back["grid"], g["grid"])
    assert back["dx"] == float("%g" % sp[0])


def test_sdf_text_cache_matches_the_reference_writer_and_reader(tmp_path):
    """tests/golden/sdf_cache_small.sdf is the file the UNMODIFIED reference wrote through
    BodySDFObj::save_SDF (BodySDFObj.cpp:105-116) for the fp32-valued grid in the .npz
    (oracle/make_golden_sdf_cache.py); `reloaded_*` is what a second reference construction read
    back through BodySDFObj::load_SDF (:86-103). The library's writer must reproduce the file byte
    for byte and its reader the reloaded fields bit for bit."""
    from assemble_them_all_b200 import contact
    z = np.load(os.path.join(GOLDEN, "sdf_cache_small.npz"))
    ref_bytes = open(os.path.join(GOLDEN, "sdf_cache_small.sdf"), "rb").read()
    ni, nj, nk = (int(v) for v in z["fresh_dims"])
    sp = z["fresh_spacing"]
    g = {"ni": ni, "nj": nj, "nk": nk, "min_box": z["fresh_min_box"], "max_box": z["fresh_max_box"], "dx": sp[0], "dy": sp[1],
         "dz": sp[2], "col_th": 0.0, "grid": z["fresh_grid"]}
    path = tmp_path / "part.sdf"
    contact.save_sdf_text(path, g)
    assert open(path, "rb").read() == ref_bytes
    back = contact.load_sdf_text(os.path.join(GOLDEN, "sdf_cache_small.sdf"), col_th=0.01)
    assert (back["ni"], back["nj"], back["nk"]) == tuple(int(v) for v in z["reloaded_dims"])
    assert np.array_equal(back["grid"], z["reloaded_grid"])
    assert np.array_equal(back["min_box"], z["reloaded_min_box"]) and np.array_equal(back["max_box"], z["reloaded_max_box"])
    assert np.array_equal([back["dx"], back["dy"], back["dz"]], z["reloaded_spacing"])
    # the round trip is lossy (6 significant digits): reloaded grids are no longer fp32-representable
    assert not np.array_equal(z["reloaded_grid"], z["fresh_grid"])
    assert not np.array_equal(z["reloaded_grid"], z["reloaded_grid"].astype(np.float32).astype(np.float64))


@pytest.mark.gpu
def test_cuda_level_set_matches_reference_grids(cuda_ctx):
    from assemble_them_all_b200 import contact
    z = _golden()
    for n in (str(x) for x in z["names"]):
        assert np.array_equal(contact.make_level_set3(cuda_ctx, *_args(z, n), exact_band=1), z[f"{n}_grid_f32"]), n
        assert np.array_equal(contact.make_level_set3(cuda_ctx, *_args(z, n), exact_band=2), z[f"{n}_grid_band2_f32"]), n


@pytest.mark.gpu
def test_cuda_precompute_sdf_full_size_matches_oracle(cuda_ctx):
    """The C1/C2 geometry (~10 000 vertices per part, dx 0.05, 10^6-cell grids): CUDA == CPU oracle
    bit for bit, and the grid is usable as an fp32-lossless SdfGrid."""
    import time
    from assemble_them_all_b200 import contact, synth
    from oracle import sdfgen_oracle as SG
    parts = synth.peg_in_hole(n_vertices=10000)
    for pid in ("0", "1"):
        V, F = parts[pid]
        V = np.float64(np.float32(V))
        t0 = time.perf_counter()
        g = contact.precompute_sdf(cuda_ctx, V, F, 0.05, col_th=0.01)
        t_gpu = time.perf_counter() - t0
        lay = SG.grid_layout(V, 0.05)
        t0 = time.perf_counter()
        ref = SG.make_level_set3(F, V.astype(np.float32), lay["min_box"], np.float32(lay["spacing"][0]),
                                 np.float32(lay["spacing"][1]), np.float32(lay["spacing"][2]), *(int(v) for v in lay["dims"]))
        t_cpu = time.perf_counter() - t0
        assert (g["ni"], g["nj"], g["nk"]) == tuple(int(v) for v in lay["dims"])
        assert np.array_equal(g["grid"].astype(np.float32), ref)
        print(f"part {pid}: {g['ni']}x{g['nj']}x{g['nk']} cells, {len(F)} triangles: GPU {t_gpu * 1e3:.1f} ms, CPU oracle {t_cpu * 1e3:.0f} ms")
        sdf = contact.SdfGrid.from_dict(cuda_ctx, g)
        assert sdf.is_fp32
        sdf.close()
