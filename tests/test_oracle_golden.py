"""CPU: the C oracle (oracle/sdfc_oracle.c) against the golden vectors the
unmodified reference produced (tests/golden/contact_small.npz)."""
import numpy as np
import pytest

from oracle import sdfc_oracle as O
from tests.helpers import assert_close


@pytest.fixture(scope="module")
def grids(golden_contact):
    O.build()
    return {n: O.Grid.from_dict(golden_contact.grid_dict(n)) for n in golden_contact.names}


def test_golden_grids_are_fp32_lossless(golden_contact):
    for n in golden_contact.names:
        g = golden_contact.grid_dict(n)
        assert g["grid"].size == g["ni"] * g["nj"] * g["nk"]
        assert np.array_equal(g["grid"].astype(np.float32).astype(np.float64), g["grid"])


@pytest.mark.parametrize("name", ["part0", "part1"])
def test_raw_queries_bit_exact(golden_contact, grids, name):
    z, g = golden_contact.z, grids[name]
    X = z[f"{name}_qX"]
    assert np.array_equal(O.query_sdf(g, X, False), z[f"{name}_q_sdf"])
    assert np.array_equal(O.query_sdf(g, X, True), z[f"{name}_q_sdf_acc"])
    assert np.array_equal(O.query_dsdf(g, X), z[f"{name}_q_dsdf"])
    assert np.array_equal(O.query_ddsdf(g, X[:1000]), z[f"{name}_q_ddsdf"])
    assert np.array_equal(O.grid_location(g, X), z[f"{name}_q_cell"])
    # SDFMesh::distance is the same sample without col_th (Mesh.cpp:143-198)
    assert np.array_equal(O.mesh_distance(g, X), z[f"{name}_q_sdf"])


@pytest.mark.parametrize("case", range(8))
def test_detection_and_forces(golden_contact, grids, case):
    G = golden_contact
    prm = G.case_params(case)
    kn = prm["kn"] * prm["scale"] * prm["contact_scale"]
    kt = prm["kt"] * prm["scale"] * prm["contact_scale"]
    for fi in range(G.n_forces):
        A, B = G.force_bodies(fi)
        E1, phi1 = G.case_pose(case, A)
        E2, phi2 = G.case_pose(case, B)
        det = O.detect(grids[B], E1, E2, G.points(A))
        assert np.array_equal(det["d"], G.out(case, fi, "d")), "per-point d must be bit-exact"
        assert np.array_equal(det["cell"], G.out(case, fi, "cell")), "cell indices must be bit-exact"
        assert np.array_equal(det["active_ids"], G.out(case, fi, "active")), "active list must be bit-exact"
        fm, Km, Dm, cnt = O.pair_force(grids[B], E1, phi1, E2, phi2, G.points(A), kn, kt, prm["mu"], prm["damping"], True)
        assert cnt[0] == len(G.out(case, fi, "active"))
        assert_close(G.out(case, fi, "fm"), fm, "fm_sub")
        assert_close(G.out(case, fi, "Km"), Km, "Km_sub")
        assert_close(G.out(case, fi, "Dm"), Dm, "Dm_sub")
        fmv, _, _, _ = O.pair_force(grids[B], E1, phi1, E2, phi2, G.points(A), kn, kt, prm["mu"], prm["damping"], False)
        assert_close(G.out(case, fi, "fm_value"), fmv, "fm_sub (value path)")
        # public wrapper: scatter into zeroed system arrays
        nm = G.ndof_m
        fs, Ks, Ds = np.zeros(nm), np.zeros((nm, nm)), np.zeros((nm, nm))
        O.scatter(nm, G.index(A), G.index(B), fm, Km, Dm, fs, Ks, Ds)
        assert_close(G.out(case, fi, "fm_sys"), fs, "fm system")
        assert_close(G.out(case, fi, "Km_sys"), Ks, "Km system")
        assert_close(G.out(case, fi, "Dm_sys"), Ds, "Dm system")


@pytest.mark.parametrize("case", [1, 5])
def test_per_contact_terms(golden_contact, grids, case):
    G = golden_contact
    prm = G.case_params(case)
    kn = prm["kn"] * prm["scale"] * prm["contact_scale"]
    kt = prm["kt"] * prm["scale"] * prm["contact_scale"]
    for fi in range(G.n_forces):
        A, B = G.force_bodies(fi)
        E1, phi1 = G.case_pose(case, A)
        E2, phi2 = G.case_pose(case, B)
        ids = G.out(case, fi, "pc_ids")
        for row, pid in enumerate(ids):
            fm, Km, Dm, _ = O.contact_term(grids[B], E1, phi1, E2, phi2, G.points(A)[pid], kn, kt, prm["mu"], prm["damping"])
            assert_close(G.out(case, fi, "pc_fm")[row], fm, f"contact {pid} fm")
            assert_close(G.out(case, fi, "pc_Km")[row], Km, f"contact {pid} Km")
            assert_close(G.out(case, fi, "pc_Dm")[row], Dm, f"contact {pid} Dm")


def test_friction_branch_mask_covers_both(golden_contact, grids):
    """Static and dynamic Coulomb branches both occur in the fixtures."""
    G = golden_contact
    n_static = n_total = 0
    for case in range(G.n_cases):
        prm = G.case_params(case)
        if prm["mu"] <= 1e-8:
            continue
        for fi in range(G.n_forces):
            A, B = G.force_bodies(fi)
            E1, phi1 = G.case_pose(case, A)
            E2, phi2 = G.case_pose(case, B)
            _, _, _, cnt = O.pair_force(grids[B], E1, phi1, E2, phi2, G.points(A), prm["kn"], prm["kt"], prm["mu"], prm["damping"], False)
            n_total += cnt[0]
            n_static += cnt[1]
    assert 0 < n_static < n_total


def test_empty_and_outside(grids):
    g = grids["part1"]
    E = np.eye(4)
    det = O.detect(g, E, E, np.zeros((0, 3)))
    assert det["d"].shape == (0,) and det["active_ids"].shape == (0,)
    far = np.array([[100.0, 0, 0], [0, -50.0, 3.0]])
    det = O.detect(g, E, E, far)
    assert np.all(det["d"] == np.float64(np.finfo(np.float32).max) + g.col_th)  # FLT_MAX outside, BodySDFObj.cpp:170
    assert len(det["active_ids"]) == 0
    fm, Km, Dm, cnt = O.pair_force(g, E, np.zeros(6), E, np.zeros(6), far, 1e6, 1e3, 0.5, 1.0, True)
    assert cnt == (0, 0) and not fm.any() and not Km.any() and not Dm.any()
