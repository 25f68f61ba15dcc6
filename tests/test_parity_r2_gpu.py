"""GPU parity tests added in round 2 (VERDICT r1 "close the parity holes"):

  * rows a6 / a7 directly: the raw FD gradient / Hessian (query_dSDF, query_ddSDF) of both device
    stencil paths against the reference's golden samples, bit for bit;
  * the batched detection kernel's 1-bit mask itself (not only its count), for every point order
    and every grid layout (fp64, fp32 cell records, fp32 bricks staged by TMA);
  * C5 at 512^3: 1e6 sampled points, mask + per-point d + raw SDFMesh.distance bit-exact against
    the oracle, for cell records (4.3 GB, size_t indexing) and for bricks;
  * the k-d point order and the brick kernel under arbitrary rotations (brick fits / does not fit);
  * sdfc_batch_detect_device on a batch whose stepper has temporal-coherence state.
"""
import os

import numpy as np
import pytest

from assemble_them_all_b200 import synth

pytestmark = pytest.mark.gpu


class _Layout:
    """SDFC_GRID_LAYOUT for the grids created inside the block."""

    def __init__(self, kind):
        self.kind = kind

    def __enter__(self):
        self.old = os.environ.get("SDFC_GRID_LAYOUT")
        if self.kind is None:
            os.environ.pop("SDFC_GRID_LAYOUT", None)
        else:
            os.environ["SDFC_GRID_LAYOUT"] = self.kind

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("SDFC_GRID_LAYOUT", None)
        else:
            os.environ["SDFC_GRID_LAYOUT"] = self.old


@pytest.mark.parametrize("layout", ["records", "bricks"])
@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("name", ["part0", "part1"])
def test_raw_gradient_and_hessian_bit_exact(cuda_ctx, golden_contact, name, fast, layout):
    """query_SDF(x, true), query_dSDF, query_ddSDF (BodySDFObj.cpp:123-205) at the golden sample points
    (inside, lattice hits, box faces, outside): the round-off-dominated FD Hessian must be reproduced
    bit for bit, it feeds Km."""
    from assemble_them_all_b200 import contact
    z = golden_contact.z
    with _Layout(layout):
        sdf = contact.SdfGrid.from_dict(cuda_ctx, golden_contact.grid_dict(name))
    assert sdf.layout == layout
    X = z[f"{name}_qX"]
    acc, g, H = sdf.stencil(X, fast=fast)
    assert np.array_equal(acc, z[f"{name}_q_sdf_acc"])
    assert np.array_equal(g, z[f"{name}_q_dsdf"])
    assert np.array_equal(H[:1000], z[f"{name}_q_ddsdf"])
    assert np.array_equal(sdf.distance(X), z[f"{name}_q_sdf"])
    sdf.close()


def _detect_all_ways(ctx, spec, pts, E1, E2, layouts=("records", "bricks"), orders=(False, True, "kd")):
    """-> oracle detection + [(layout, order, mask, n_active, d_per_point)] from the batched kernel."""
    from assemble_them_all_b200 import contact
    from oracle import sdfc_oracle as O
    og = O.Grid.from_spec(spec)
    ref = O.detect(og, E1, E2, pts)
    ref_mask = np.zeros(len(pts), dtype=np.uint8)
    ref_mask[ref["active_ids"]] = 1
    out = []
    for layout in layouts:
        with _Layout(layout):
            sdf = contact.SdfGrid.from_spec(ctx, spec)
        assert sdf.layout == layout
        for order in orders:
            cp = contact.ContactPoints(ctx, pts, morton_sort=order)
            pair = contact.ForceGeneralSDFContact(ctx, cp, sdf, 1e6, 1e3)
            batch = contact.ContactBatch(ctx, 1, 2, [pair], [0], [1])
            E = np.stack([E1, E2])[None]
            _, _, _, nact = batch.eval(E, np.zeros((1, 2, 6)), deriv=False)
            mask = batch.active_mask(0, 0)
            q = pair.query(E1, E2, want_cell=True)
            out.append((layout, order, mask, int(nact[0, 0]), q))
            batch.close(); pair.close(); cp.close()
        sdf.close()
    return ref, ref_mask, out


@pytest.mark.parametrize("rot", [0.01, 0.6, 2.2])
def test_batched_mask_bit_exact_all_layouts_and_orders(cuda_ctx, rot):
    """The mask words k_detect_mask / k_detect_brick write, expanded to original point order, against
    the oracle's active set: a 96^3 grid (bricks forced) with 60 000 points, small and large relative
    rotations (large ones make some tiles' bricks exceed shared memory: both branches of the brick
    kernel run), points partly outside the grid box."""
    rng = np.random.default_rng(7)
    spec = synth.sphere_grid(96, col_th=0.03)
    pts = rng.uniform(-5.6, 5.6, size=(60000, 3))
    E1 = synth.make_E(synth.axis_angle_rotation([0.3, -0.5, 0.8], rot), [0.11, -0.07, 0.05])
    E2 = synth.make_E(synth.axis_angle_rotation([-0.2, 0.4, 0.1], 0.3 * rot), [-0.02, 0.03, 0.01])
    ref, ref_mask, got = _detect_all_ways(cuda_ctx, spec, pts, E1, E2)
    assert 0 < ref_mask.sum() < len(pts)
    for layout, order, mask, nact, q in got:
        what = f"layout {layout}, order {order}"
        assert np.array_equal(mask, ref_mask), what
        assert nact == int(ref_mask.sum()), what
        assert np.array_equal(q["d"], ref["d"]) and np.array_equal(q["cell"], ref["cell"]), what
        assert np.array_equal(q["active"], ref_mask), what


def test_c5_512_grid_million_points_bit_exact(cuda_ctx):
    """C5 at its largest grid: 512^3 (cell records 4.3 GB / bricks 0.55 GB), 1e6 uniform points.
    Mask of the batched kernel, per-point d / cell and raw SDFMesh.distance against the oracle."""
    from assemble_them_all_b200 import contact, scenes
    from oracle import sdfc_oracle as O
    spec, pts = scenes.sweep_scene(1_000_000, 512)
    R = synth.axis_angle_rotation([0.3, -0.2, 0.9], 0.01)
    E1, E2 = synth.make_E(R, [0.01, -0.02, 0.015]), np.eye(4)
    ref, ref_mask, got = _detect_all_ways(cuda_ctx, spec, pts, E1, E2, orders=(False, "kd"))
    frac = ref_mask.mean()
    assert 0.02 < frac < 0.10
    for layout, order, mask, nact, q in got:
        what = f"layout {layout}, order {order}"
        assert np.array_equal(mask, ref_mask), what
        assert nact == int(ref_mask.sum()), what
        assert np.array_equal(q["d"], ref["d"]) and np.array_equal(q["cell"], ref["cell"]), what
    # raw sampling (SDFMesh.distance, Mesh.cpp:143-189) through the bricked layout
    og = O.Grid.from_spec(spec)
    sdf = contact.SdfGrid.from_spec(cuda_ctx, spec)
    assert sdf.layout.startswith("bricks")
    X = pts[:200000]
    assert np.array_equal(sdf.distance(X), O.mesh_distance(og, X))
    sdf.close()


def test_kd_order_tiles_are_compact_and_forces_agree(cuda_ctx, golden_contact):
    """k-d order: same contacts, forces within 1e-9 of the reference goldens (the summation order
    changes, nothing else)."""
    from assemble_them_all_b200 import contact
    from tests.helpers import assert_close
    G = golden_contact
    sdf = {n: contact.SdfGrid.from_dict(cuda_ctx, G.grid_dict(n)) for n in G.names}
    pts = {n: contact.ContactPoints(cuda_ctx, G.points(n), morton_sort="kd") for n in G.names}
    for case in (0, 3, 5):
        prm = G.case_params(case)
        for fi in range(G.n_forces):
            A, B = G.force_bodies(fi)
            E1, phi1 = G.case_pose(case, A)
            E2, phi2 = G.case_pose(case, B)
            p = contact.ForceGeneralSDFContact(cuda_ctx, pts[A], sdf[B], prm["kn"], prm["kt"], prm["mu"], prm["damping"])
            p.set_scale(prm["scale"])
            p.set_contact_scale(prm["contact_scale"])
            fm, Km, Dm, n = p.force_deriv_sub(E1, phi1, E2, phi2)
            assert n == len(G.out(case, fi, "active"))
            assert_close(G.out(case, fi, "fm"), fm, f"kd case {case} force {fi} fm")
            assert_close(G.out(case, fi, "Km"), Km, f"kd case {case} force {fi} Km")
            assert_close(G.out(case, fi, "Dm"), Dm, f"kd case {case} force {fi} Dm")
            p.close()


def test_detect_device_ignores_stale_stepper_culling_state(cuda_ctx):
    """sdfc_batch_detect_device on a batch whose stepper keeps temporal-coherence margins must
    not read them. Step a rollout batch (margins and pair-skip modes are now those of the last stepper
    evaluation), then detect at unrelated poses: counts and masks equal those of a fresh batch."""
    import torch
    from assemble_them_all_b200 import scenes
    rs = scenes.c2_planner_rollouts(n_envs=8, n_vertices=2000)
    sim, batch, pairs, sdfs, pts = rs.upload(cuda_ctx)
    sim.reset(rs.q0, rs.qdot0)
    for b in range(len(rs.bodies)):
        sim.set_body_external_force(b, rs.f_ext[:, b])
    sim.forward(30)
    assert batch.skin_stats()["rebuilds"] > 0
    rs2 = scenes.c2_bfs_rollouts(n_envs=8, n_vertices=2000)   # same geometry, contact-rich poses
    assert [len(b.points) for b in rs2.bodies] == [len(b.points) for b in rs.bodies]
    dev = torch.device("cuda", 0)
    dE = torch.from_numpy(np.ascontiguousarray(np.swapaxes(rs2.E, -1, -2))).to(dev)
    dna = torch.zeros((8, 2), dtype=torch.int32, device=dev)
    batch.detect_device(dE.data_ptr(), dna.data_ptr())
    cuda_ctx.synchronize()
    got = dna.cpu().numpy()
    masks = [[batch.active_mask(e, p) for p in range(2)] for e in range(8)]
    fresh, *_ = rs2.upload(cuda_ctx)
    _, _, _, want = fresh.eval(rs2.E, rs2.phi, deriv=False)
    assert want.sum() > 0
    assert np.array_equal(got, want)
    for e in range(8):
        for p in range(2):
            assert np.array_equal(masks[e][p], fresh.active_mask(e, p))
    sim.close()
