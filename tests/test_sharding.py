"""CPU: the multi-GPU host logic (rollout partition + final gather of outcome
records) with torch.distributed/gloo, world_size 2 (and 3), no GPU involved."""
import os
import socket

import numpy as np
import pytest

from assemble_them_all_b200 import sharding


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 5, 1024, 1031):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _fake_outcomes(ids):
    """Deterministic stand-in for what a rank's GPU rollouts produce."""
    ids = np.asarray(ids)
    status = ((ids * 7919) % 13 == 5).astype(np.float64)
    q = np.stack([np.sin(ids * 0.37 + k) for k in range(3)], axis=1)
    return sharding.make_records(ids, status, 100 + ids % 7, ids % 6, ids % 2, q)


def _worker(rank, world, port, n_total, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(n_total, rank, world)
    table = sharding.gather_records(_fake_outcomes(np.arange(lo, hi)), n_total, dist)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), table)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world,n_total", [(2, 1024), (2, 7), (3, 100)])
def test_gather_matches_single_process(tmp_path, world, n_total):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    expect = _fake_outcomes(np.arange(n_total))
    single = sharding.gather_records(expect, n_total, None)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert np.array_equal(got, expect), f"rank {r}"
    assert sharding.first_success(single) == sharding.first_success(np.load(tmp_path / "rank0.npy"))
    if n_total >= 100:
        assert sharding.first_success(single) >= 0


def test_gather_rejects_misordered_tables():
    rec = _fake_outcomes(np.arange(4))[::-1].copy()
    with pytest.raises(RuntimeError):
        sharding.gather_records(rec, 4, None)
