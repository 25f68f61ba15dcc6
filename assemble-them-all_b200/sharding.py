"""Multi-GPU partitioning of planner rollouts.

Rollouts (one Simulation each: assembly, move part, BFS node, action) never
interact, so they are block-partitioned over the ranks of one node — one process
per GPU, no data-path collective, no per-step exchange. The only communication is
one all_gather of fixed-size outcome records per planning stage, which replaces
the reference's per-process result queue (utils/parallel.py:17-66): rank order =
rollout order, so the merged list is in the reference's queue order and "first
success in order" decisions (run_joint_plan.py:539-575) are reproducible.

Backend: torch.distributed — NCCL over NVLink on the GPU box, gloo in CPU tests.
"""
from __future__ import annotations

import numpy as np

# outcome record of one rollout: [rollout id, status, steps done, action id, touched, q_final (up to 6)]
RECORD_WIDTH = 5 + 6


def shard_bounds(n_items, rank, world):
    """Contiguous block partition; the first n_items % world ranks take one extra item."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def make_records(ids, status, steps, action, touched, q_final):
    """Pack per-rollout outcomes into the float64 record layout the gather moves."""
    ids = np.asarray(ids)
    rec = np.zeros((len(ids), RECORD_WIDTH))
    rec[:, 0] = ids
    rec[:, 1] = status
    rec[:, 2] = steps
    rec[:, 3] = action
    rec[:, 4] = touched
    q = np.asarray(q_final, dtype=np.float64).reshape(len(ids), -1)
    rec[:, 5:5 + q.shape[1]] = q
    return rec


def gather_records(local_records, n_total, dist=None, device=None):
    """All ranks -> the full (n_total, RECORD_WIDTH) table in rollout order.
    `dist` is torch.distributed (initialised) or None for a single process."""
    local_records = np.ascontiguousarray(local_records, dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = local_records
    else:
        import torch
        world, rank = dist.get_world_size(), dist.get_rank()
        cap = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
        buf = torch.zeros((cap, RECORD_WIDTH), dtype=torch.float64, device=device)
        buf[: len(local_records)] = torch.from_numpy(local_records).to(buf.device)
        parts = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)   # the single collective of a planning stage
        rows = []
        for r, p in enumerate(parts):
            lo, hi = shard_bounds(n_total, r, world)
            rows.append(p[: hi - lo].cpu().numpy())
        out = np.concatenate(rows, axis=0)
    if len(out) != n_total or not np.array_equal(out[:, 0], np.arange(n_total)):
        raise RuntimeError("gather_records: merged table is not in rollout order")
    return out


def first_success(records):
    """Index of the first successful rollout in reference order, or -1 (status 1 = disassembled)."""
    hits = np.nonzero(records[:, 1] == 1)[0]
    return int(hits[0]) if len(hits) else -1
