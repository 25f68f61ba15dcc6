"""CPU: the synthetic input generators are deterministic and well-formed."""
import numpy as np

from assemble_them_all_b200 import synth


def test_peg_in_hole_is_watertight_and_sized():
    parts = synth.peg_in_hole(n_vertices=2000)
    for V, F in parts.values():
        assert synth.mesh_is_watertight(V, F)
        assert abs(len(V) - 2000) / 2000 < 0.15
    Vp, _ = parts["0"]
    Vb, _ = parts["1"]
    assert np.allclose(Vp.min(0), [-1, -1, -2]) and np.allclose(Vp.max(0), [1, 1, 4])
    assert np.allclose(Vb.min(0), [-3, -3, -2]) and np.allclose(Vb.max(0), [3, 3, 2])
    # volume by divergence theorem: block minus hole
    def vol(V, F):
        a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
        return np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0
    assert np.isclose(vol(*parts["0"]), 2 * 2 * 6)
    assert np.isclose(vol(*parts["1"]), 6 * 6 * 4 - 2.1 * 2.1 * 4)


def test_generators_are_deterministic():
    a, b = synth.peg_in_hole(600), synth.peg_in_hole(600)
    for k in a:
        assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1])
    g1, g2 = synth.sphere_grid(16), synth.sphere_grid(16)
    assert np.array_equal(g1.grid, g2.grid)
    assert np.array_equal(g1.grid.astype(np.float32).astype(np.float64), g1.grid)


def test_peg_array_has_n_parts():
    parts = synth.peg_array(nx=2, ny=2, n_vertices=300, n_parts=5)
    assert len(parts) == 5
    for V, F in parts.values():
        assert synth.mesh_is_watertight(V, F)


def test_planner_xml_mentions_both_directions():
    xml = synth.planner_xml("/tmp/x", "0", ["1", "2"])
    assert xml.count("<general_SDF_contact general_body=") == 4
    assert 'timestep="1e-3"' in xml and 'col_th="0.01"' in xml


def test_c4_scene_bodies_and_candidate_layout():
    """C4 scene (scenes.c4_bodies / c4_candidate_rollouts): analytic grids agree with the meshes
    (every vertex lies on the zero level set up to the grid's interpolation error), mass properties
    follow from the volumes, and a candidate's contact forces come in the planner XML's order."""
    from assemble_them_all_b200 import scenes
    from oracle import sdfc_oracle as O
    O.build()
    bodies = scenes.c4_bodies(nx=2, ny=2, n_parts=5, n_vertices=500, sdf_dx=0.1)
    assert [b.name for b in bodies] == [f"part{i}" for i in range(5)]
    for b in bodies:
        g = O.Grid.from_dict({"ni": b.grid.ni, "nj": b.grid.nj, "nk": b.grid.nk, "min_box": b.grid.min_box,
                              "max_box": b.grid.max_box, "dx": b.grid.dx, "dy": b.grid.dy, "dz": b.grid.dz,
                              "col_th": 0.0, "grid": b.grid.grid})
        assert np.abs(O.query_sdf(g, b.points)).max() < 0.1
        assert np.array_equal(b.grid.grid.astype(np.float32).astype(np.float64), b.grid.grid)
    s = 10.0 / 5.0
    assert np.isclose(bodies[1].mass, (1.0 * 1.0 * 4.0) * s ** 3)
    assert np.isclose(bodies[0].mass, (5.0 * 5.0 * 2.0 - 4 * 1.1 * 1.1 * 2.0) * s ** 3)
    rs = scenes.c4_candidate_rollouts(bodies, 2)
    assert [b.name for b in rs.bodies] == ["part2", "part0", "part1", "part3", "part4"]
    assert [b.joint_type for b in rs.bodies] == ["translational"] + ["fixed"] * 4
    assert rs.pairs == [(0, 1), (1, 0), (0, 2), (2, 0), (0, 3), (3, 0), (0, 4), (4, 0)]
    assert rs.n_envs == 6 and np.array_equal(rs.f_ext[4, 0], [0, 0, 0, 0, 0, 100.0]) and not rs.f_ext[:, 1:].any()
