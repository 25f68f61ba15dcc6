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
