"""CPU: host logic of the batched BFS planner (assemble_them_all_b200.planner).

* goal-test geometry (HullPair) on configurations with known answers;
* vectorised similarity rules against the scalar restatement of run_joint_plan.py:427-455;
* the batched planner, run over a pool of UNMODIFIED reference simulators (one per
  environment, oracle/_ref/ref_probe), must return exactly what the sequential planner
  loop returns on one reference simulator: status, action sequence, the whole substep
  path bit for bit, and the number of rollouts consumed in reference order.
The same comparison against the CUDA stepper is tests/test_planner_gpu.py."""
import numpy as np
import pytest

from assemble_them_all_b200 import planner as PL
from assemble_them_all_b200 import synth
from tests.drivers import seq_planner as SP


def _E(R=np.eye(3), p=(0, 0, 0)):
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = p
    return E


def _cube(c, s):
    c, s = np.asarray(c, float), np.asarray(s, float) / 2
    return np.array([[c[0] + i * s[0], c[1] + j * s[1], c[2] + k * s[2]] for i in (-1, 1) for j in (-1, 1) for k in (-1, 1)])


def test_hull_dedup_on_subdivided_box():
    V, _ = synth.box_mesh((2.0, 1.0, 3.0), h=0.25)
    h = PL.Hull(V)
    assert len(h.vertices) == 8 and len(h.normals) == 3 and len(h.edges) == 3 and len(h.planes) == 6


def test_hull_pair_known_configurations():
    still = np.vstack([_cube((0, 0, 0), (4, 4, 4))])
    hp = PL.HullPair(_cube((0, 0, 0), (1, 1, 1)), still)
    poses = np.array([
        _E(p=(0, 0, 0)),        # inside: no surface contact, still contains move -> not disassembled
        _E(p=(2.0, 0, 0)),      # straddles the face -> surface contact
        _E(p=(2.6, 0, 0)),      # clear by 0.1
        _E(p=(2.4, 2.4, 2.4)),  # overlapping the corner
        _E(p=(2.51, 2.51, 2.51)),
    ])
    inter, _ = hp.solids_intersect(poses)
    assert inter.tolist() == [True, True, False, True, False]
    surf, _ = hp.surfaces_intersect(poses)
    assert surf.tolist() == [False, True, False, True, False]
    assert hp.is_disassembled(poses).tolist() == [False, False, True, False, True]


def test_hull_pair_edge_edge_axis():
    """Two cubes rotated 45 degrees about different axes, separated only by an edge-edge axis."""
    c = np.cos(np.pi / 4)
    Rz = np.array([[c, -c, 0], [c, c, 0], [0, 0, 1.0]])
    Rx = np.array([[1.0, 0, 0], [0, c, -c], [0, c, c]])
    still = _cube((0, 0, 0), (2, 2, 2)) @ Rx.T
    hp = PL.HullPair(_cube((0, 0, 0), (2, 2, 2)), still)
    # rotated cubes have a vertical/horizontal edge pointing at each other along x = y... use brute force as the judge
    rng = np.random.default_rng(3)
    poses, truth = [], []
    for _ in range(60):
        p = rng.uniform(-3, 3, 3)
        poses.append(_E(Rz, p))
        # brute force: sample the moved cube densely and test against the still hull's planes
        g = np.linspace(-1, 1, 41)
        P = np.array(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1).T @ Rz.T + p
        inside = (P @ hp.still.planes[:, :3].T + hp.still.planes[:, 3] <= 1e-12).all(axis=1)
        truth.append(bool(inside.any()))
    inter, _ = hp.solids_intersect(np.array(poses))
    # the sampling can miss grazing overlaps but never reports a false overlap
    assert not np.any(np.array(truth) & ~inter)
    assert np.sum(inter != np.array(truth)) <= 3
    assert inter.any() and (~inter).any()


def test_similarity_matches_scalar_form():
    rng = np.random.default_rng(11)
    for rotation in (False, True):
        nd = 6 if rotation else 3
        hits = 0
        for _ in range(200):
            path = rng.normal(scale=0.05, size=(rng.integers(0, 12), nd))
            if rotation:
                path[:, 3:] = rng.normal(scale=0.6, size=(len(path), 3))
            new = rng.normal(scale=0.03, size=nd)
            if rotation:
                new[3:] = rng.normal(scale=0.6, size=3)
            a = PL.any_state_similar(path, new, rotation)
            b = SP.any_state_similar(list(path), new, rotation)
            assert a == b
            hits += a
        assert 0 < hits < 200


def test_action_to_force():
    assert np.array_equal(PL.action_to_force([0, 0, -1], 100.0), [0, 0, 0, 0, 0, -100.0])
    assert np.array_equal(PL.action_to_force([0, 1, 0, 0, 0, 0], 100.0), [0, 300.0, 0, 0, 0, 0])


# ---------------------------------------------------------------- planner vs sequential loop
def _ref_scene(parts, joint_type, dx):
    from oracle import ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    make = lambda: ref_scene.probe_from_assembly(parts, "0", ["1"], joint_type, sdf_dx=dx)  # noqa: E731
    p = make()
    hulls = PL.HullPair(np.array(p.get_body_vertices("part0")).T, np.array(p.get_body_vertices("part1", True)).T)
    return make, p, hulls


@pytest.mark.parametrize("n_envs", [6, 15])
def test_batched_bfs_equals_sequential_depth2(n_envs):
    parts = synth.l_channel()
    make, p, hulls = _ref_scene(parts, "translational", 0.1)
    seq = SP.SequentialBFSPlanner(p, "part0", hulls)
    st, path, acts, stages = seq.plan(PL.TRANS_ACTIONS, max_depth=8)
    assert st == "Success" and acts == [4, 0]  # +x along the tunnel, then +z out of the shaft
    pool = SP.RefSimPool(make, n_envs, "part0", 3, ["part0", "part1"])
    bp = PL.BatchedBFSPlanner(pool, 0, hulls)
    r = bp.plan(max_depth=8)
    assert r.status == st and r.actions == acts and r.stages == stages
    assert r.rollouts == seq.rollouts
    assert np.array_equal(r.path, path)
    assert r.rollouts_run >= r.rollouts and r.env_steps >= seq.env_steps


def test_batched_bfs_failure_at_max_depth_and_action_order():
    R = synth.axis_angle_rotation(np.array([0.0, 1.0, 0.0]), -np.pi / 2)  # hole axis z -> -x
    parts = synth.reorient(synth.peg_in_hole(n_vertices=300, peg_z=(-1.0, 4.0)), R)
    make, p, hulls = _ref_scene(parts, "translational", 0.2)
    seq = SP.SequentialBFSPlanner(p, "part0", hulls)
    st, path, acts, stages = seq.plan(PL.TRANS_ACTIONS, max_depth=1)
    pool = SP.RefSimPool(make, 6, "part0", 3, ["part0", "part1"])
    bp = PL.BatchedBFSPlanner(pool, 0, hulls)
    r = bp.plan(max_depth=1)
    assert (r.status, r.actions, r.stages, r.rollouts) == (st, acts, stages, seq.rollouts)
    assert st == "Success" and len(acts) == 1 and acts[0] in (4, 5)
    assert np.array_equal(r.path, path)


def test_batched_bfs_rotational_level():
    parts = synth.l_channel()
    make, p, hulls = _ref_scene(parts, "free3d-exp", 0.1)
    seq = SP.SequentialBFSPlanner(p, "part0", hulls, rotation=True)
    st, path, acts, stages = seq.plan(PL.ROT_ACTIONS, max_depth=1)
    pool = SP.RefSimPool(make, 12, "part0", 6, ["part0", "part1"])
    bp = PL.BatchedBFSPlanner(pool, 0, hulls, rotation=True)
    r = bp.plan(max_depth=1)
    assert (r.status, r.actions, r.stages, r.rollouts) == (st, acts, stages, seq.rollouts)
    assert st == "Failure" and r.rollouts == 12


# ---------------------------------------------------------------- sequence planning
def _seq_setup(n_envs_factory):
    """Three-part scene; returns plan_path callables for the sequential and the batched arm."""
    from oracle import ref_scene
    if not ref_scene.available():
        pytest.skip("oracle/_ref not built")
    # tunnel along +z, shaft along +y: the cube's plan is [+z, +y] = the first child of the root,
    # so max_depth 2 (node expansions, run_joint_plan.py:574-576) reaches it
    parts = synth.reorient(synth.l_channel_with_plug(), np.array([[0.0, 1, 0], [0, 0, 1], [1, 0, 0]]))
    cache = {}

    def scene(move_id, still_ids):
        key = (move_id, tuple(still_ids))
        if key not in cache:
            sub = {k: parts[k] for k in [move_id, *still_ids]}
            make = lambda: ref_scene.probe_from_assembly(sub, move_id, list(still_ids), "translational", sdf_dx=0.1)  # noqa: E731
            p = make()
            names = [f"part{i}" for i in [move_id, *still_ids]]
            still_v = np.vstack([np.array(p.get_body_vertices(n, True)).T for n in names[1:]])
            hulls = PL.HullPair(np.array(p.get_body_vertices(names[0])).T, still_v)
            cache[key] = (make, p, hulls, names)
        return cache[key]

    def plan_seq(move_id, still_ids, rotation, max_depth):
        _, p, hulls, names = scene(move_id, still_ids)
        return SP.SequentialBFSPlanner(p, names[0], hulls).plan(PL.TRANS_ACTIONS, max_depth=max_depth)[0]

    def plan_batched(move_id, still_ids, rotation, max_depth):
        make, _, hulls, names = scene(move_id, still_ids)
        sim = n_envs_factory(make, names)
        return PL.BatchedBFSPlanner(sim, 0, hulls).plan(max_depth=max_depth).status

    return plan_seq, plan_batched


def test_sequence_planner_matches_reference_loop():
    plan_seq, plan_batched = _seq_setup(lambda make, names: SP.RefSimPool(make, 6, names[0], 3, names))
    for seed in (1, 2):
        st, seq, count, trace = SP.reference_prog_queue(["0", "1", "2"], plan_seq, False, seed, max_trials=12)
        r = PL.SequencePlanner(["0", "1", "2"], plan_batched).plan_sequence("prog-queue", False, seed, max_trials=12)
        assert (r.status, r.sequence, r.trials) == (st, seq, count)
        assert [(m, d, s) for m, d, _, s in r.log] == trace
        assert st == "Success" and seq.index("2") < seq.index("0")   # the plug leaves before the cube
        assert any(s == "Failure" for _, _, s in trace)
        assert ("0", 2, "Success") in trace                          # the cube needs the deeper search


def test_sequence_planner_variants_consume_rng_like_the_reference():
    """queue / random variants with a scripted path planner: same trial order as a literal walk."""
    outcomes = {"a": ["Failure", "Success"], "b": ["Success"], "c": ["Failure", "Failure", "Success"]}

    def scripted():
        seen = {k: 0 for k in outcomes}

        def plan(move_id, still_ids, rotation, max_depth):
            i = min(seen[move_id], len(outcomes[move_id]) - 1)
            seen[move_id] += 1
            return outcomes[move_id][i]
        return plan

    for seed in (0, 3):
        r = PL.SequencePlanner(["a", "b", "c"], scripted()).plan_sequence("queue", False, seed)
        np.random.seed(seed)
        q = ["a", "b", "c"]
        np.random.shuffle(q)
        assert [m for m, *_ in r.log][:3] == q
        assert r.status == "Success" and len(r.sequence) == 2
        r2 = PL.SequencePlanner(["a", "b", "c"], scripted()).plan_sequence("random", False, seed)
        np.random.seed(seed)
        first = str(np.random.choice(["a", "b", "c"]))
        assert r2.log[0][0] == first
