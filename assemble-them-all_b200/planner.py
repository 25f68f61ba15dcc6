"""Batched BFS disassembly planner on top of the device-resident stepper.

Host-side mirror of the reference's `BFSPlanner` (examples/run_joint_plan.py:421-684):
same actions, same queue order, same stop rules, same returned (status, path). What
changes is the execution: every (node, action) rollout of a stage is one environment
of a `BatchedSimulation`, all of them advance `frame_skip` BDF1 steps per call, and
rollouts that hit their stop rule are retired with `set_active` while the others
continue. The outcome is resolved in the reference's sequential order (nodes in queue
order, actions in list order), so the plan is the one the sequential loop returns.

Geometry of the goal test. The reference asks trimesh / python-fcl (not vendored, not in
this image) whether the convex hull of the moving part still touches the
hull of the still parts (run_joint_plan.py:291-302). `HullPair` restates that predicate
with numpy + scipy.spatial.ConvexHull: the hull SURFACES intersect iff the solids
intersect (separating-axis test over face normals and edge-edge cross products) and
neither solid contains the other; without surface contact the reference's bounding-box
containment rule is applied verbatim. Both planner arms of the parity tests (this
batched one and the sequential driver over the stock redmax_py) use this predicate.

The simulation backend only needs the `BatchedSimulation` surface used here
(n_envs, ndof_r, reset, set_state, get_state, set_body_external_force, set_active,
forward(trajectory=True), get_body_E0i), so the host logic is testable on CPU against a
pool of stock simulators.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.spatial import ConvexHull
from scipy.spatial.transform import Rotation

TRANS_ACTIONS = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0]], dtype=np.float64)
# run_joint_plan.py:585-598: the six translations first (q order is (p, r), the action is
# (rotation, translation)), then the six rotations
ROT_ACTIONS = np.array([
    [0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 0, -1], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, -1, 0], [0, 0, 0, 1, 0, 0],
    [0, 0, 0, -1, 0, 0], [0, 0, 1, 0, 0, 0], [0, 0, -1, 0, 0, 0], [0, 1, 0, 0, 0, 0], [0, -1, 0, 0, 0, 0],
    [1, 0, 0, 0, 0, 0], [-1, 0, 0, 0, 0, 0]], dtype=np.float64)


def action_to_force(action, force_mag):
    """PhysicsPlanner.apply_action (run_joint_plan.py:270-278): unit action * force_mag;
    3-vectors are translations, 6-vectors are (rotation * 3, translation)."""
    a = np.asarray(action, dtype=np.float64)
    a = a / np.linalg.norm(a) * force_mag
    if len(a) == 3:
        return np.concatenate([np.zeros(3), a])
    return np.concatenate([a[:3] * 3, a[3:]])


# ------------------------------------------------------------------ goal-test geometry
def _unique_directions(d, tol=1e-7):
    """Unit directions up to sign, deduplicated."""
    d = d[np.linalg.norm(d, axis=1) > 1e-12]
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    # canonical sign: first component of magnitude > tol is positive
    big = np.abs(d) > 1e-9
    first = np.argmax(big, axis=1)
    sign = np.sign(d[np.arange(len(d)), first])
    d = d * sign[:, None]
    key = np.round(d / tol).astype(np.int64)
    _, idx = np.unique(key, axis=0, return_index=True)
    return d[np.sort(idx)]


class Hull:
    """Convex hull of a vertex cloud: hull vertices, outward face planes, edge directions."""

    def __init__(self, points):
        points = np.asarray(points, dtype=np.float64)
        h = ConvexHull(points)
        self.vertices = points[h.vertices]
        eq = h.equations  # outward normal n, offset: n.x + off <= 0 inside
        key = np.round(eq / 1e-7).astype(np.int64)
        _, idx = np.unique(key, axis=0, return_index=True)
        self.planes = eq[np.sort(idx)]
        self.normals = _unique_directions(self.planes[:, :3])
        # edges shared by two non-coplanar facets
        tri = h.simplices
        fn = eq[:, :3]
        edge_faces = {}
        for f, (a, b, c) in enumerate(tri):
            for u, v in ((a, b), (b, c), (c, a)):
                edge_faces.setdefault((min(u, v), max(u, v)), []).append(f)
        dirs = []
        for (u, v), fs in edge_faces.items():
            if len(fs) == 2 and np.dot(fn[fs[0]], fn[fs[1]]) > 1.0 - 1e-10:
                continue  # interior edge of a coplanar facet fan
            dirs.append(points[v] - points[u])
        self.edges = _unique_directions(np.array(dirs)) if dirs else np.zeros((0, 3))


def _overlap_on_axes(axes, VA, VB, tol):
    """axes (..., m, 3), VA (..., na, 3), VB (nb, 3) -> True where no axis separates."""
    pa = np.einsum("...md,...nd->...mn", axes, VA)
    pb = np.einsum("...md,nd->...mn", axes, VB)
    sep = (pa.min(-1) > pb.max(-1) + tol) | (pb.min(-1) > pa.max(-1) + tol)
    return ~sep.any(-1)


class HullPair:
    """is_disassembled of the reference (run_joint_plan.py:291-302) for a batch of poses."""

    def __init__(self, vertices_move_body, vertices_still_world, tol=1e-9):
        self.move = Hull(vertices_move_body)
        self.still = Hull(vertices_still_world)
        vs = np.asarray(vertices_still_world, dtype=np.float64)
        self.min_box_still, self.max_box_still = vs.min(axis=0), vs.max(axis=0)
        self.tol = tol

    def solids_intersect(self, E):
        E = np.asarray(E, dtype=np.float64).reshape(-1, 4, 4)
        R, p = E[:, :3, :3], E[:, :3, 3]
        VA = np.einsum("eij,nj->eni", R, self.move.vertices) + p[:, None, :]
        nA = np.einsum("eij,mj->emi", R, self.move.normals)
        eA = np.einsum("eij,mj->emi", R, self.move.edges)
        ok = _overlap_on_axes(nA, VA, self.still.vertices, self.tol)
        nB = np.broadcast_to(self.still.normals, (len(E),) + self.still.normals.shape)
        ok &= _overlap_on_axes(nB, VA, self.still.vertices, self.tol)
        if len(self.move.edges) and len(self.still.edges):
            cr = np.cross(eA[:, :, None, :], self.still.edges[None, None, :, :]).reshape(len(E), -1, 3)
            nrm = np.linalg.norm(cr, axis=-1, keepdims=True)
            cr = np.where(nrm > 1e-9, cr / np.maximum(nrm, 1e-300), 0.0)  # parallel edges: null axis never separates
            ok &= _overlap_on_axes(cr, VA, self.still.vertices, self.tol)
        return ok, VA

    def surfaces_intersect(self, E):
        """What the FCL triangle-mesh test of the two hulls reports: solids overlap and
        neither hull lies strictly inside the other."""
        inter, VA = self.solids_intersect(E)
        pl = self.still.planes
        a_in_b = (np.einsum("end,md->enm", VA, pl[:, :3]) + pl[:, 3] < -self.tol).all(axis=(1, 2))
        E = np.asarray(E, dtype=np.float64).reshape(-1, 4, 4)
        # still vertices in the moving hull's frame
        loc = np.einsum("eji,enj->eni", E[:, :3, :3], self.still.vertices[None] - E[:, None, :3, 3])
        pm = self.move.planes
        b_in_a = (np.einsum("end,md->enm", loc, pm[:, :3]) + pm[:, 3] < -self.tol).all(axis=(1, 2))
        return inter & ~a_in_b & ~b_in_a, VA

    def is_disassembled(self, E):
        has_collision, VA = self.surfaces_intersect(E)
        mn, mx = VA.min(axis=1), VA.max(axis=1)
        move_contain_still = (mn <= self.min_box_still).all(-1) & (mx >= self.max_box_still).all(-1)
        still_contain_move = (self.min_box_still <= mn).all(-1) & (self.max_box_still >= mx).all(-1)
        return ~has_collision & ~(move_contain_still | still_contain_move)


# ------------------------------------------------------------------ similarity rules
TRANS_DIST_TH = 0.05   # BFSPlanner.trans_dist_th (run_joint_plan.py:424)
QUAT_DIST_TH = 0.5     # BFSPlanner.quat_dist_th (run_joint_plan.py:425)


def _quat_wxyz(q_rot):
    """BFSPlanner.get_quat (run_joint_plan.py:430-431): the rotation dofs are read as
    'xyz' Euler angles, exactly as the reference does."""
    xyzw = Rotation.from_euler("xyz", q_rot).as_quat()
    return xyzw[..., [3, 0, 1, 2]]


def quat_distance(q0, q1):
    """pyquaternion 0.9.9 Quaternion.distance(q0, q1) = |log(q0^-1 q1)| for unit
    quaternions (w, x, y, z), vectorised over the leading axis of q1."""
    w0, v0 = q0[..., 0], -q0[..., 1:]
    w1, v1 = q1[..., 0], q1[..., 1:]
    w = w0 * w1 - np.sum(v0 * v1, axis=-1)
    v = w0[..., None] * v1 + w1[..., None] * v0 + np.cross(v0, v1)
    n = np.sqrt(w * w + np.sum(v * v, axis=-1))
    vn = np.linalg.norm(v, axis=-1)
    ang = np.where(vn > 1e-14, np.arccos(np.clip(w / n, -1.0, 1.0)), 0.0)
    return np.sqrt(np.log(n) ** 2 + ang ** 2)


def any_state_similar(path, new_q, rotation):
    """BFSPlanner.any_state_similar (run_joint_plan.py:451-455) over a (k, ndof) path."""
    if len(path) == 0:
        return False
    path = np.asarray(path)
    trans = np.linalg.norm(path[:, :3] - new_q[:3], axis=1) < TRANS_DIST_TH
    if not rotation:
        return bool(trans.any())
    qd = quat_distance(_quat_wxyz(new_q[3:])[None, :], _quat_wxyz(path[:, 3:])) < QUAT_DIST_TH
    return bool((trans & qd).any())


# ------------------------------------------------------------------ the planner
@dataclass
class RolloutResult:
    node: int
    action: int
    status: str          # "Success" | "Similar"
    path: np.ndarray     # (k, ndof): every substep q of this rollout (without the parent path)
    chunks: int


@dataclass
class PlanResult:
    status: str                      # "Success" | "Failure" | "Start with goal"
    path: np.ndarray | None          # (k, ndof) substep states from the start to the goal
    actions: list = field(default_factory=list)   # action index per BFS level on the way to the goal
    stages: int = 0
    rollouts: int = 0                # rollouts whose outcome the reference order consumed
    rollouts_run: int = 0            # rollouts executed on the device (speculative ones included)
    env_steps: int = 0
    forward_calls: int = 0


class BatchedBFSPlanner:
    """BFSPlanner.plan_trans / plan_rot (run_joint_plan.py:511-684), batched.

    sim: BatchedSimulation-like backend; move_body: index of the moving link;
    hulls: HullPair; q_init: joint state after Simulation::reset (zeros in the planner XML)."""

    def __init__(self, sim, move_body, hulls, rotation=False, force_mag=100.0, frame_skip=100, q_init=None,
                 max_chunks=1000):
        self.sim, self.move_body, self.hulls = sim, move_body, hulls
        self.rotation = bool(rotation)
        self.ndof = 6 if rotation else 3
        if sim.ndof_r != self.ndof:
            raise ValueError(f"planner expects {self.ndof} reduced dofs, simulation has {sim.ndof_r}")
        self.force_mag, self.frame_skip = float(force_mag), int(frame_skip)
        self.actions = ROT_ACTIONS if rotation else TRANS_ACTIONS
        self.forces = np.array([action_to_force(a, self.force_mag) for a in self.actions])
        self.q_init = np.zeros(self.ndof) if q_init is None else np.asarray(q_init, dtype=np.float64)
        self.max_chunks = int(max_chunks)
        self.env_steps = 0
        self.forward_calls = 0
        self.rollouts_run = 0

    # -- one wave: up to n_envs (node state, action) rollouts, run to their stop rule
    def _run_wave(self, jobs):
        """jobs: list of (q_start (ndof,), parent_path list/array, action index).
        Returns per job (status, rollout_path (k, ndof), chunks)."""
        sim, E, fs = self.sim, self.sim.n_envs, self.frame_skip
        n = len(jobs)
        assert 0 < n <= E
        q = np.tile(self.q_init, (E, 1))
        f = np.zeros((E, 6))
        for e, (q0, _, a) in enumerate(jobs):
            q[e] = q0
            f[e] = self.forces[a]
        sim.reset(q, np.zeros_like(q))                       # sim.reset(); set_state(state)
        sim.set_body_external_force(self.move_body, f)       # apply_action(action)
        active = np.zeros(E, dtype=bool)
        active[:n] = True
        paths = [[] for _ in range(n)]
        status = [None] * n
        chunks = np.zeros(n, dtype=np.int64)
        while active.any():
            # self.set_state(self.get_state()): keep q, zero qdot (run_joint_plan.py:262-265,549)
            qc, _ = sim.get_state()
            sim.set_state(qc, np.zeros_like(qc))
            sim.set_active(active)
            Q, _ = sim.forward(fs, trajectory=True)
            self.forward_calls += 1
            self.env_steps += int(active.sum()) * fs
            E0 = sim.get_body_E0i()[:, self.move_body]
            ids = np.nonzero(active)[0]
            done = self.hulls.is_disassembled(E0[ids])
            for k, e in enumerate(ids):
                paths[e].append(Q[:, e, :].copy())
                chunks[e] += 1
                if done[k]:
                    status[e], active[e] = "Success", False
                    continue
                parent = jobs[e][1]
                hist = paths[e][:-1]
                prev = np.concatenate([np.asarray(parent).reshape(-1, self.ndof)] + hist) if (len(parent) or hist) \
                    else np.zeros((0, self.ndof))
                if any_state_similar(prev, Q[-1, e, :], self.rotation):
                    status[e], active[e] = "Similar", False   # back and forth (run_joint_plan.py:566-567)
                elif chunks[e] >= self.max_chunks:
                    raise RuntimeError("rollout exceeded max_chunks without reaching a stop rule")
        sim.set_active(None)
        self.rollouts_run += n
        return [(status[e], np.concatenate(paths[e]), int(chunks[e])) for e in range(n)]

    def _run_jobs(self, jobs):
        out = []
        for i in range(0, len(jobs), self.sim.n_envs):
            out += self._run_wave(jobs[i:i + self.sim.n_envs])
        return out

    def plan(self, max_depth=None, nodes_per_wave=None):
        """Returns PlanResult. max_depth follows the reference: node expansions for the
        translational planner (run_joint_plan.py:574-576), BFS levels for the rotational
        one (:678-681). Wall-clock timeouts are not modelled (max_time = inf)."""
        self.env_steps = self.forward_calls = self.rollouts_run = 0
        # Simulation::reset state; "Start with goal" is PhysicsPlanner.plan's rule only
        # (run_joint_plan.py:335-337); BFS starts expanding right away.
        na = len(self.actions)
        # queue entries: (q (ndof,), path (k, ndof), action list)
        queue = [(self.q_init.copy(), np.zeros((0, self.ndof)), [])]
        stages = consumed = 0
        if nodes_per_wave is None:
            nodes_per_wave = max(1, self.sim.n_envs // na)
        while True:
            if self.rotation:
                batch_nodes = list(queue)          # the whole level (run_joint_plan.py:617-619)
                queue = []
            else:
                batch_nodes = queue[:nodes_per_wave]  # speculative: the next nodes in FIFO order
            if max_depth is not None and not self.rotation:
                batch_nodes = batch_nodes[:max_depth - stages]
            jobs = [(q, path, a) for (q, path, _) in batch_nodes for a in range(na)]
            res = self._run_jobs(jobs)
            stage_states = []
            for ni, (q, path, acts) in enumerate(batch_nodes):
                for a in range(na):
                    st, rpath, _ = res[ni * na + a]
                    consumed += 1
                    full = np.concatenate([path, rpath])
                    if st == "Success":
                        return PlanResult("Success", full, acts + [a], stages, consumed, self.rollouts_run,
                                          self.env_steps, self.forward_calls)
                    entry = (rpath[-1].copy(), full, acts + [a])
                    if self.rotation:
                        stage_states.append(entry)
                    else:
                        queue.append(entry)
                if not self.rotation:
                    queue.pop(0)
                    stages += 1
                    if max_depth is not None and stages == max_depth:
                        return PlanResult("Failure", None, [], stages, consumed, self.rollouts_run, self.env_steps,
                                          self.forward_calls)
            if self.rotation:
                # sorted(stage_states, key=-len(path)) is stable (run_joint_plan.py:678)
                queue = sorted(stage_states, key=lambda s: -len(s[1]))
                stages += 1
                if max_depth is not None and stages == max_depth:
                    return PlanResult("Failure", None, [], stages, consumed, self.rollouts_run, self.env_steps,
                                      self.forward_calls)


# ------------------------------------------------------------------ sequence planning
SUCCESS_STATUS = ("Success", "Start with goal")   # run_multi_plan.py:36-38
FAILURE_STATUS = ("Timeout", "Failure")


@dataclass
class SequenceResult:
    status: str                 # "Success" | "Failure"
    sequence: list              # part ids in removal order
    trials: int                 # seq_count
    log: list = field(default_factory=list)   # (move_id, max_depth, rotation tried, status) per trial


class SequencePlanner:
    """The sequence planners of examples/run_multi_plan.py on top of a path-planner factory.

    plan_path(move_id, still_ids, rotation, max_depth) -> status string; with the batched BFS
    planner every call is one stage-parallel plan on the device. The trial order reproduces
    the reference's use of numpy's legacy global RNG (seed, shuffle / choice, one randint per
    trial), so a given seed visits the candidate parts in the same order.
    variant: "random" (:69-133), "queue" (:136-225), "prog-queue" (:228-330)."""

    def __init__(self, part_ids, plan_path):
        self.part_ids = [str(p) for p in part_ids]
        assert len(self.part_ids) > 1
        self.plan_path = plan_path
        n = len(self.part_ids)
        self.max_seq_count = (1 + n) * n // 2 - 1

    def _attempt(self, move_id, nodes, rotation, max_depth, log):
        still = [p for p in nodes if p != move_id]
        status = self.plan_path(move_id, still, False, max_depth)
        tried_rot = False
        if status in FAILURE_STATUS and rotation:
            status = self.plan_path(move_id, still, True, max_depth)
            tried_rot = True
        log.append((move_id, max_depth, tried_rot, status))
        return status

    def plan_sequence(self, variant="prog-queue", rotation=False, seed=1, max_trials=None):
        np.random.seed(seed)
        nodes = list(self.part_ids)
        sequence, log = [], []
        seq_count = 0
        status = "Failure"
        if variant == "random":
            while seq_count < self.max_seq_count:
                move_id = str(np.random.choice(nodes))
                np.random.randint(self.max_seq_count)      # curr_seed
                st = self._attempt(move_id, nodes, rotation, None, log)
                seq_count += 1
                if st in SUCCESS_STATUS:
                    nodes.remove(move_id)
                    sequence.append(move_id)
                if len(nodes) == 1:
                    status = "Success"
                    break
            return SequenceResult(status, sequence, seq_count, log)
        if variant == "queue":
            active = list(nodes)
            np.random.shuffle(active)
            last_active, inactive = list(active), []
            while seq_count < self.max_seq_count:
                move_id = active.pop(0)
                np.random.randint(self.max_seq_count)
                st = self._attempt(move_id, nodes, rotation, None, log)
                seq_count += 1
                if st in SUCCESS_STATUS:
                    nodes.remove(move_id)
                    sequence.append(move_id)
                else:
                    inactive.append(move_id)
                if len(nodes) == 1:
                    status = "Success"
                    break
                if len(active) == 0:
                    active, inactive = list(inactive), []
                    if active == last_active:
                        break
                    last_active = list(active)
            return SequenceResult(status, sequence, seq_count, log)
        if variant != "prog-queue":
            raise ValueError(f"unknown sequence planner {variant}")
        active = [(p, 1) for p in nodes]
        np.random.shuffle(active)
        inactive = []
        while True:
            move_id, max_depth = active.pop(0)
            np.random.randint(self.max_seq_count)
            st = self._attempt(move_id, nodes, rotation, max_depth, log)
            seq_count += 1
            if st in SUCCESS_STATUS:
                nodes.remove(move_id)
                sequence.append(move_id)
            else:
                inactive.append((move_id, max_depth + 1))
            if len(nodes) == 1:
                status = "Success"
                break
            if len(active) == 0:
                active, inactive = list(inactive), []
            if max_trials is not None and seq_count >= max_trials:   # stands in for seq_max_time (:321-323)
                break
        return SequenceResult(status, sequence, seq_count, log)
