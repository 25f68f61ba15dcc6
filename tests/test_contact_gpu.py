"""GPU parity tests: the CUDA path (through the C-ABI) against the golden
vectors of the unmodified reference and against the CPU oracle.
Bar: integers / masks / per-point distances bit-exact; fm, Km, Dm within 1e-9
relative (scaled by the largest magnitude of the compared array)."""
import numpy as np
import pytest

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(cuda_ctx, golden_contact):
    from assemble_them_all_b200 import contact
    G = golden_contact
    sdf = {n: contact.SdfGrid.from_dict(cuda_ctx, G.grid_dict(n)) for n in G.names}
    pts = {n: contact.ContactPoints(cuda_ctx, G.points(n), morton_sort=True) for n in G.names}
    pts_unsorted = {n: contact.ContactPoints(cuda_ctx, G.points(n), morton_sort=False) for n in G.names}
    return {"sdf": sdf, "pts": pts, "pts_unsorted": pts_unsorted}


def make_pair(ctx, scene, G, fi, prm, sorted_pts=True):
    from assemble_them_all_b200 import contact
    A, B = G.force_bodies(fi)
    p = contact.ForceGeneralSDFContact(ctx, scene["pts" if sorted_pts else "pts_unsorted"][A], scene["sdf"][B],
                                       prm["kn"], prm["kt"], prm["mu"], prm["damping"],
                                       index_contact=G.index(A), index_sdf=G.index(B))
    p.set_scale(prm["scale"])
    p.set_contact_scale(prm["contact_scale"])
    return p


@pytest.mark.parametrize("b", [0.05000000074505806, 0.1, float(np.float32(10.0 / 511)), 0.0333333333333, 1.0, 0.7853981633974483])
def test_exact_division_matches_ieee(cuda_ctx, b):
    """(x - min) / dx is evaluated from a precomputed reciprocal; it must equal the
    IEEE quotient for every numerator (cell index and lerp weights depend on it)."""
    assert cuda_ctx.selftest_exact_div(200_000_000, 12345, b) == 0


def test_grids_use_fp32_lossless_storage(scene):
    assert all(s.is_fp32 for s in scene["sdf"].values())


@pytest.mark.parametrize("name", ["part0", "part1"])
def test_sdf_distance_bit_exact(scene, golden_contact, name):
    z = golden_contact.z
    X = z[f"{name}_qX"]
    got = scene["sdf"][name].distance(X)
    assert np.array_equal(got, z[f"{name}_q_sdf"])
    assert scene["sdf"][name].min_distance(X) == z[f"{name}_q_sdf"].min()


@pytest.mark.parametrize("sorted_pts", [True, False])
@pytest.mark.parametrize("case", range(8))
def test_pair_against_reference_golden(cuda_ctx, scene, golden_contact, case, sorted_pts):
    G = golden_contact
    prm = G.case_params(case)
    for fi in range(G.n_forces):
        A, B = G.force_bodies(fi)
        E1, phi1 = G.case_pose(case, A)
        E2, phi2 = G.case_pose(case, B)
        pair = make_pair(cuda_ctx, scene, G, fi, prm, sorted_pts)
        q = pair.query(E1, E2)
        assert np.array_equal(q["d"], G.out(case, fi, "d")), "per-point d must be bit-exact"
        assert np.array_equal(q["cell"], G.out(case, fi, "cell")), "cell indices must be bit-exact"
        ref_active = np.zeros(len(q["d"]), dtype=np.uint8)
        ref_active[G.out(case, fi, "active")] = 1
        assert np.array_equal(q["active"], ref_active), "active mask must be bit-exact"
        assert q["n_active"] == len(G.out(case, fi, "active"))

        fm, Km, Dm, n = pair.force_deriv_sub(E1, phi1, E2, phi2)
        assert n == len(G.out(case, fi, "active"))
        assert_close(G.out(case, fi, "fm"), fm, f"case {case} force {fi} fm_sub")
        assert_close(G.out(case, fi, "Km"), Km, f"case {case} force {fi} Km_sub")
        assert_close(G.out(case, fi, "Dm"), Dm, f"case {case} force {fi} Dm_sub")
        fmv, n2 = pair.force_sub(E1, phi1, E2, phi2)
        assert n2 == n
        assert_close(G.out(case, fi, "fm_value"), fmv, f"case {case} force {fi} fm_sub (value)")

        # public-wrapper semantics: accumulate into caller arrays
        nm = G.ndof_m
        fs, Ks, Ds = np.ones(nm), np.ones((nm, nm)), np.ones((nm, nm))
        pair.computeForceWithDerivative(fs, Ks, Ds, E1, phi1, E2, phi2)
        assert_close(G.out(case, fi, "fm_sys") + 1, fs, "fm system (+=)")
        assert_close(G.out(case, fi, "Km_sys") + 1, Ks, "Km system (+=)")
        assert_close(G.out(case, fi, "Dm_sys") + 1, Ds, "Dm system (+=)")
        assert pair.touched == (n > 0)
        pair.close()


def test_results_are_reproducible(cuda_ctx, scene, golden_contact):
    """Fixed-order reductions: two evaluations are bit-identical."""
    G = golden_contact
    prm = G.case_params(5)
    A, B = G.force_bodies(0)
    E1, phi1 = G.case_pose(5, A)
    E2, phi2 = G.case_pose(5, B)
    pair = make_pair(cuda_ctx, scene, G, 0, prm)
    a = pair.force_deriv_sub(E1, phi1, E2, phi2)
    b = pair.force_deriv_sub(E1, phi1, E2, phi2)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    pair.close()


@pytest.mark.parametrize("cases", [(0, 1), (4,), (5,), (3, ), (7,)])
def test_batch_matches_golden_system_arrays(cuda_ctx, scene, golden_contact, cases):
    """E environments sharing parameters: system arrays = sum over forces of the
    reference's public-wrapper outputs (Robot.cpp:546-547 loop)."""
    from assemble_them_all_b200 import contact
    G = golden_contact
    prm = G.case_params(cases[0])
    for c in cases:
        assert G.case_params(c) == prm
    pairs = [make_pair(cuda_ctx, scene, G, fi, prm) for fi in range(G.n_forces)]
    body_id = {n: i for i, n in enumerate(G.names)}
    assert all(G.index(n) == 6 * body_id[n] for n in G.names)
    a = [body_id[G.force_bodies(fi)[0]] for fi in range(G.n_forces)]
    b = [body_id[G.force_bodies(fi)[1]] for fi in range(G.n_forces)]
    reps = 3
    envs = [c for c in cases for _ in range(reps)]
    batch = contact.ContactBatch(cuda_ctx, len(envs), len(G.names), pairs, a, b)
    E = np.stack([np.stack([G.case_pose(c, n)[0] for n in G.names]) for c in envs])
    phi = np.stack([np.stack([G.case_pose(c, n)[1] for n in G.names]) for c in envs])
    fm, Km, Dm, nact = batch.eval(E, phi, deriv=True)
    fmv, _, _, nact2 = batch.eval(E, phi, deriv=False)
    assert np.array_equal(nact, nact2)
    for e, c in enumerate(envs):
        ref_f = sum(G.out(c, fi, "fm_sys") for fi in range(G.n_forces))
        ref_K = sum(G.out(c, fi, "Km_sys") for fi in range(G.n_forces))
        ref_D = sum(G.out(c, fi, "Dm_sys") for fi in range(G.n_forces))
        assert_close(ref_f, fm[e], f"env {e} fm")
        assert_close(ref_K, Km[e], f"env {e} Km")
        assert_close(ref_D, Dm[e], f"env {e} Dm")
        assert_close(ref_f, fmv[e], f"env {e} fm (value)")
        assert [int(v) for v in nact[e]] == [len(G.out(c, fi, "active")) for fi in range(G.n_forces)]
    # identical environments give bit-identical rows
    assert np.array_equal(fm[0], fm[1]) and np.array_equal(Km[0], Km[2])
    batch.close()
    for p in pairs:
        p.close()


def _oracle_scene(rng, n_pts, grid_n, col_th):
    from assemble_them_all_b200 import synth
    spec = synth.sphere_grid(grid_n, length=10.0, radius_frac=0.35, col_th=col_th)
    # points on a shell around the sphere surface (radius 3.5), some outside the box
    dirs = rng.normal(size=(n_pts, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    pts = dirs * rng.uniform(3.2, 3.9, size=(n_pts, 1))
    pts[: n_pts // 50] *= 3.0  # far outside the grid box
    return spec, pts


@pytest.mark.parametrize("mu,damping", [(0.0, 0.0), (0.6, 0.0), (0.0, 20.0), (0.4, 15.0)])
def test_against_oracle_on_analytic_sphere(cuda_ctx, mu, damping):
    """Larger seeded case checked against the CPU oracle (itself pinned to the reference)."""
    from assemble_them_all_b200 import contact, synth
    from oracle import sdfc_oracle as O
    rng = np.random.default_rng(99)
    spec, pts = _oracle_scene(rng, 6000, 64, col_th=0.02)
    og = O.Grid.from_spec(spec)
    sdf = contact.SdfGrid.from_spec(cuda_ctx, spec)
    cp = contact.ContactPoints(cuda_ctx, pts)
    pair = contact.ForceGeneralSDFContact(cuda_ctx, cp, sdf, 1e5, 1e3, mu, damping)
    for trial in range(3):
        E1 = synth.make_E(synth.random_rotation(rng), rng.uniform(-0.3, 0.3, 3))
        E2 = synth.make_E(synth.random_rotation(rng), rng.uniform(-0.3, 0.3, 3))
        phi1, phi2 = rng.normal(size=6), rng.normal(size=6)
        det = O.detect(og, E1, E2, pts)
        q = pair.query(E1, E2)
        assert np.array_equal(q["d"], det["d"]) and np.array_equal(q["cell"], det["cell"])
        assert np.array_equal(np.nonzero(q["active"])[0], det["active_ids"])
        assert len(det["active_ids"]) > 500
        fm, Km, Dm, cnt = O.pair_force(og, E1, phi1, E2, phi2, pts, 1e5, 1e3, mu, damping, True)
        gfm, gKm, gDm, n = pair.force_deriv_sub(E1, phi1, E2, phi2)
        assert n == cnt[0]
        assert_close(fm, gfm, "fm"); assert_close(Km, gKm, "Km"); assert_close(Dm, gDm, "Dm")
        fmv, _, _, _ = O.pair_force(og, E1, phi1, E2, phi2, pts, 1e5, 1e3, mu, damping, False)
        gfmv, _ = pair.force_sub(E1, phi1, E2, phi2)
        assert_close(fmv, gfmv, "fm (value)")
    pair.close(); cp.close(); sdf.close()


@pytest.mark.parametrize("knobs", [{"SDFC_FORCE_G": "4"}, {"SDFC_FORCE_CAP": "4"}, {"SDFC_FORCE_CAP": "1", "SDFC_FORCE_G": "3"}])
@pytest.mark.parametrize("mu,damping", [(0.0, 0.0), (0.5, 10.0)])
def test_force_work_list_with_grouped_rounds(cuda_ctx, monkeypatch, knobs, mu, damping):
    """The force kernels' work list when one item holds several 32-contact rounds: G forced up by a larger
    minimum (SDFC_FORCE_G) or by a list capacity the evaluation overflows (SDFC_FORCE_CAP, read when the
    batch is created). Same contacts, other grouping of the partial sums: the oracle bar, not bit equality."""
    from assemble_them_all_b200 import contact, synth
    from oracle import sdfc_oracle as O
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(7)
    spec, pts = _oracle_scene(rng, 9000, 64, col_th=0.02)
    og = O.Grid.from_spec(spec)
    sdf = contact.SdfGrid.from_spec(cuda_ctx, spec)
    cp = contact.ContactPoints(cuda_ctx, pts)
    pair = contact.ForceGeneralSDFContact(cuda_ctx, cp, sdf, 1e5, 1e3, mu, damping)
    for trial in range(2):
        E1 = synth.make_E(synth.random_rotation(rng), rng.uniform(-0.3, 0.3, 3))
        E2 = synth.make_E(synth.random_rotation(rng), rng.uniform(-0.3, 0.3, 3))
        phi1, phi2 = rng.normal(size=6), rng.normal(size=6)
        fm, Km, Dm, cnt = O.pair_force(og, E1, phi1, E2, phi2, pts, 1e5, 1e3, mu, damping, True)
        assert cnt[0] > 700   # more than 20 rounds
        gfm, gKm, gDm, n = pair.force_deriv_sub(E1, phi1, E2, phi2)
        assert n == cnt[0]
        assert_close(fm, gfm, "fm"); assert_close(Km, gKm, "Km"); assert_close(Dm, gDm, "Dm")
        fmv, _, _, _ = O.pair_force(og, E1, phi1, E2, phi2, pts, 1e5, 1e3, mu, damping, False)
        gfmv, _ = pair.force_sub(E1, phi1, E2, phi2)
        assert_close(fmv, gfmv, "fm (value)")
    pair.close(); cp.close(); sdf.close()


def test_fp64_grid_storage_path(cuda_ctx):
    """Grids that are not fp32-representable (reloaded .sdf text caches,
    BodySDFObj.cpp:86-103) stay fp64 and still match the oracle bit for bit."""
    from assemble_them_all_b200 import contact, synth
    from oracle import sdfc_oracle as O
    rng = np.random.default_rng(5)
    spec = synth.sphere_grid(40, col_th=0.03, dtype_round=None)
    og = O.Grid.from_spec(spec)
    sdf = contact.SdfGrid.from_spec(cuda_ctx, spec)
    assert not sdf.is_fp32
    X = rng.uniform(-5.5, 5.5, size=(20000, 3))
    assert np.array_equal(sdf.distance(X), O.mesh_distance(og, X))
    _, pts = _oracle_scene(rng, 3000, 40, 0.03)
    cp = contact.ContactPoints(cuda_ctx, pts)
    pair = contact.ForceGeneralSDFContact(cuda_ctx, cp, sdf, 1e5, 1e3, 0.3, 4.0)
    E1 = synth.make_E(synth.random_rotation(rng), rng.uniform(-0.2, 0.2, 3))
    E2 = synth.make_E(synth.random_rotation(rng), rng.uniform(-0.2, 0.2, 3))
    phi1, phi2 = rng.normal(size=6), rng.normal(size=6)
    det = O.detect(og, E1, E2, pts)
    q = pair.query(E1, E2)
    assert np.array_equal(q["d"], det["d"]) and np.array_equal(q["cell"], det["cell"])
    fm, Km, Dm, _ = O.pair_force(og, E1, phi1, E2, phi2, pts, 1e5, 1e3, 0.3, 4.0, True)
    gfm, gKm, gDm, _ = pair.force_deriv_sub(E1, phi1, E2, phi2)
    assert_close(fm, gfm, "fm"); assert_close(Km, gKm, "Km"); assert_close(Dm, gDm, "Dm")


def test_edge_cases(cuda_ctx, scene, golden_contact):
    from assemble_them_all_b200 import contact
    G = golden_contact
    sdf = scene["sdf"]["part1"]
    # empty point set
    empty = contact.ContactPoints(cuda_ctx, np.zeros((0, 3)))
    pair = contact.ForceGeneralSDFContact(cuda_ctx, empty, sdf, 1e6, 1e3)
    q = pair.query(np.eye(4), np.eye(4))
    assert q["n_active"] == 0 and q["d"].shape == (0,)
    # all points far outside: FLT_MAX + col_th, no contacts, zero blocks
    far = contact.ContactPoints(cuda_ctx, np.array([[100.0, 0, 0], [0, -50.0, 3.0], [7, 7, 7]]))
    pair2 = contact.ForceGeneralSDFContact(cuda_ctx, far, sdf, 1e6, 1e3, 0.5, 1.0)
    q = pair2.query(np.eye(4), np.eye(4))
    assert np.all(q["d"] == np.float64(np.finfo(np.float32).max) + sdf.col_th) and q["n_active"] == 0
    fm, Km, Dm, n = pair2.force_deriv_sub(np.eye(4), np.zeros(6), np.eye(4), np.zeros(6))
    assert n == 0 and not fm.any() and not Km.any() and not Dm.any()
    # ragged: point count not a multiple of 32, every point active (deep inside part1's solid)
    rng = np.random.default_rng(3)
    inside = np.column_stack([rng.uniform(1.5, 2.5, 77), rng.uniform(-2.5, 2.5, 77), rng.uniform(-1.5, 1.5, 77)])
    cp = contact.ContactPoints(cuda_ctx, inside)
    pair3 = contact.ForceGeneralSDFContact(cuda_ctx, cp, sdf, 1e6, 1e3)
    E2 = G.z["part1_E0"]
    q = pair3.query(E2, E2)  # contact body posed like the SDF body: points are body-frame coordinates of part1
    assert q["n_active"] == 77 and q["active"].all()
    # wrong argument types fail loudly, in the reference's words
    with pytest.raises(contact.SdfContactError):
        contact.ForceGeneralSDFContact(cuda_ctx, cp, object())
    with pytest.raises(contact.SdfContactError):
        contact.SdfGrid(cuda_ctx, 4, 4, 4, [0, 0, 0], [1, 1, 1], 0.3, 0.3, 0.3, 0.0, np.zeros(10))


def test_results_do_not_depend_on_stale_device_memory():
    """SDFC_POISON=1 fills every fresh device allocation of the library with 0xFF bytes (NaN / -1). A batch
    evaluation (both variants), planner rollouts of both joint types and a free-fall batch must give the
    same bytes with and without it: nothing is read before the library wrote it, and creation-time
    null-stream uploads are ordered before the first launch on the (non-blocking) context stream - the
    r1i plug-in flakiness was exactly that race."""
    import os
    import subprocess
    import sys
    driver = os.path.join(os.path.dirname(os.path.abspath(__file__)), "drivers", "poison_probe.py")
    digests = []
    for poison in ("0", "1", "1"):
        r = subprocess.run([sys.executable, driver], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, SDFC_POISON=poison))
        assert r.returncode == 0, r.stdout + r.stderr
        digests.append([ln for ln in r.stdout.splitlines() if ln.startswith("DIGEST")][-1])
    assert digests[0] == digests[1] == digests[2]
