"""Sequential BFS planner over a `redmax_py.Simulation`-like object: the checker arm of
the planner parity tests. It walks the loops of the reference's BFSPlanner
(examples/run_joint_plan.py:511-684) one rollout and one `forward(1)` at a time, with the
similarity rules written in the reference's scalar form (:427-455). The goal test needs
trimesh + python-fcl in the reference (not in this image); the stand-in
is the product's `HullPair` predicate, shared by both arms.

Test infrastructure only - nothing in the product imports this file.
"""
import math

import numpy as np
from scipy.spatial.transform import Rotation

TRANS_TH, QUAT_TH = 0.05, 0.5  # run_joint_plan.py:424-425


def _quat(state):  # run_joint_plan.py:430-431
    x, y, z, w = Rotation.from_euler("xyz", state[3:]).as_quat()
    return np.array([w, x, y, z])


def _quat_distance(q0, q1):
    """pyquaternion Quaternion.distance: norm of log(q0^-1 * q1), scalar form."""
    w0, x0, y0, z0 = q0[0], -q0[1], -q0[2], -q0[3]   # inverse of a unit quaternion
    w1, x1, y1, z1 = q1
    w = w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1
    x = w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1
    y = w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1
    z = w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1
    n = math.sqrt(w * w + x * x + y * y + z * z)
    vn = math.sqrt(x * x + y * y + z * z)
    ang = math.acos(max(-1.0, min(1.0, w / n))) if vn > 1e-14 else 0.0
    return math.sqrt(math.log(n) ** 2 + ang ** 2)


def state_similar(state, new_state, rotation):
    trans = np.linalg.norm(state[:3] - new_state[:3])
    if not rotation:
        return trans < TRANS_TH
    return trans < TRANS_TH and _quat_distance(_quat(new_state), _quat(state)) < QUAT_TH


def any_state_similar(path, new_state, rotation):
    for state in path:
        if state_similar(state, new_state, rotation):
            return True
    return False


class SequentialBFSPlanner:
    def __init__(self, sim, move_name, hulls, rotation=False, force_mag=100.0, frame_skip=100):
        self.sim, self.move_name, self.hulls = sim, move_name, hulls
        self.rotation, self.force_mag, self.frame_skip = rotation, force_mag, frame_skip
        self.env_steps = 0
        self.rollouts = 0

    def _q(self):
        return np.array(self.sim.get_joint_q(self.move_name), dtype=np.float64)

    def _set_state(self, q):
        self.sim.set_joint_q(self.move_name, q)
        self.sim.zero_joint_qdot(self.move_name)

    def _apply(self, action):
        a = np.asarray(action, dtype=np.float64)
        a = a / np.linalg.norm(a) * self.force_mag
        f = np.concatenate([np.zeros(3), a]) if len(a) == 3 else np.concatenate([a[:3] * 3, a[3:]])
        self.sim.set_body_external_force(self.move_name, f)

    def _is_disassembled(self):
        E = np.array(self.sim.get_body_E0i(self.move_name), dtype=np.float64)
        return bool(self.hulls.is_disassembled(E[None])[0])

    def _rollout(self, q_start, curr_path, action):
        """-> (status, new_q, temp_path)"""
        temp_path = list(curr_path)
        self.sim.reset()
        self._set_state(q_start)
        self._apply(action)
        self.rollouts += 1
        while True:
            self._set_state(self._q())
            for _ in range(self.frame_skip):
                self.sim.forward(1)
                new_q = self._q()
                temp_path.append(new_q)
                self.env_steps += 1
            if self._is_disassembled():
                return "Success", new_q, temp_path
            if any_state_similar(temp_path[:-self.frame_skip], new_q, self.rotation):
                return "Similar", new_q, temp_path

    def plan(self, actions, max_depth=None):
        """-> (status, path or None, action index list, stages)"""
        self.sim.reset()
        states = [(self._q(), [], [])]
        n_stages = 0
        while True:
            if not self.rotation:
                q, curr_path, acts = states.pop(0)
                for ai, action in enumerate(actions):
                    st, new_q, temp_path = self._rollout(q, curr_path, action)
                    if st == "Success":
                        return "Success", np.array(temp_path), acts + [ai], n_stages
                    states.append((new_q, temp_path, acts + [ai]))
            else:
                stage_states = []
                while len(states) > 0:
                    q, curr_path, acts = states.pop(0)
                    for ai, action in enumerate(actions):
                        st, new_q, temp_path = self._rollout(q, curr_path, action)
                        if st == "Success":
                            return "Success", np.array(temp_path), acts + [ai], n_stages
                        stage_states.append((new_q, temp_path, acts + [ai]))
                states = sorted(stage_states, key=lambda x: -len(x[1]))
            n_stages += 1
            if n_stages == max_depth:
                return "Failure", None, [], n_stages


class RefSimPool:
    """BatchedSimulation-shaped backend made of stock simulators, one per environment:
    lets the batched planner's host logic run on CPU against the sequential arm."""

    def __init__(self, make_sim, n_envs, move_name, ndof, body_names):
        self.sims = [make_sim() for _ in range(n_envs)]
        self.n_envs, self.ndof_r, self.move_name = n_envs, ndof, move_name
        self.body_names = body_names
        self.active = np.ones(n_envs, dtype=bool)
        self.n_bodies = len(body_names)

    def reset(self, q=None, qdot=None):
        for e, s in enumerate(self.sims):
            s.reset()
            s.set_joint_q(self.move_name, np.zeros(self.ndof_r) if q is None else np.asarray(q)[e])
            s.zero_joint_qdot(self.move_name)

    def set_state(self, q, qdot):
        assert not np.any(qdot)
        for e, s in enumerate(self.sims):
            s.set_joint_q(self.move_name, np.asarray(q)[e])
            s.zero_joint_qdot(self.move_name)

    def get_state(self):
        q = np.array([s.get_joint_q(self.move_name) for s in self.sims])
        qd = np.array([s.get_joint_qdot(self.move_name) for s in self.sims])
        return q, qd

    def set_body_external_force(self, body, force):
        force = np.broadcast_to(force, (self.n_envs, 6))
        for e, s in enumerate(self.sims):
            s.set_body_external_force(self.body_names[body], np.ascontiguousarray(force[e]))

    def set_active(self, active=None):
        self.active = np.ones(self.n_envs, dtype=bool) if active is None else np.asarray(active, dtype=bool).copy()

    def forward(self, n_steps, trajectory=False):
        Q = np.zeros((n_steps, self.n_envs, self.ndof_r))
        QD = np.zeros_like(Q)
        for e, s in enumerate(self.sims):
            if not self.active[e]:
                continue
            for t in range(n_steps):
                s.forward(1)
                Q[t, e] = s.get_joint_q(self.move_name)
                QD[t, e] = s.get_joint_qdot(self.move_name)
        return (Q, QD) if trajectory else None

    def get_body_E0i(self):
        return np.array([[np.array(s.get_body_E0i(n)) for n in self.body_names] for s in self.sims])


def reference_prog_queue(part_ids, plan_path, rotation, seed, max_trials=None):
    """ProgressiveQueueSequencePlanner.plan_sequence (examples/run_multi_plan.py:252-330) walked
    literally; plan_path(move_id, still_ids, rotation, max_depth) -> status."""
    n = len(part_ids)
    max_seq_count = (1 + n) * n // 2 - 1
    np.random.seed(seed)
    nodes = list(part_ids)
    seq_status, sequence, seq_count = "Failure", [], 0
    active_queue = [(node, 1) for node in nodes]
    np.random.shuffle(active_queue)
    inactive_queue = []
    trace = []
    while True:
        all_ids = list(nodes)
        move_id, max_depth = active_queue.pop(0)
        still_ids = all_ids.copy()
        still_ids.remove(move_id)
        np.random.randint(max_seq_count)
        status = plan_path(move_id, still_ids, False, max_depth)
        if status in ("Timeout", "Failure") and rotation:
            status = plan_path(move_id, still_ids, True, max_depth)
        seq_count += 1
        trace.append((move_id, max_depth, status))
        if status in ("Success", "Start with goal"):
            nodes.remove(move_id)
            sequence.append(move_id)
        else:
            inactive_queue.append([move_id, max_depth + 1])
        if len(nodes) == 1:
            seq_status = "Success"
            break
        if len(active_queue) == 0:
            active_queue = inactive_queue.copy()
            inactive_queue = []
        if max_trials is not None and seq_count >= max_trials:
            break
    return seq_status, sequence, seq_count, trace
