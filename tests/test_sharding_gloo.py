"""world_size-2 gloo test of the multi-GPU host logic (trial sharding + the statistics all-reduce)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from twoace_b200 import parallel as par


def test_shard_ranges_partition():
    for total in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_results(lo, hi):
    """Deterministic per-trial 'results' keyed by the global trial id (what a rank would compute)."""
    t = np.arange(lo, hi)
    rng_vals = np.sin(t * 12.9898) * 0.5 + 0.5
    info = np.zeros((hi - lo, 16))
    info[:, 0] = rng_vals
    info[:, 2] = (t % 3 == 0)
    info[:, 3] = (t % 7 == 0)
    info[:, 15] = 1000 + t
    cells = t % 4
    mse = rng_vals * 0.1
    mse[t % 11 == 5] = np.nan      # failed solves are skipped, not summed
    met = np.stack([mse, 2.0 + rng_vals, 2.5 + rng_vals, 0.9 * rng_vals], axis=1)   # twoace_metrics_batch layout
    met[np.isnan(mse)] = np.nan
    ang = np.stack([3.0 + rng_vals] * 6, axis=1)                                    # twoace_angle_metrics_batch layout
    return cells, mse, info, met, ang


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = par.shard_range(total, rank, world)
    cells, mse, info, met, ang = _fake_results(lo, hi)
    s = par.all_reduce_stats(par.local_stats(cells, 4, mse, info, met, ang))
    q.put((rank, s))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stats_equal_single_process():
    total, world = 101, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cells, mse, info, met, ang = _fake_results(0, total)
    ref = par.local_stats(cells, 4, mse, info, met, ang)
    assert ref.shape == (4, len(par.METRICS)) and np.all(ref[:, 6:] > 0)      # the gains / projection error ride along
    ok = np.isfinite(mse)
    assert abs(ref[:, 7].sum() - met[ok, 2].sum()) < 1e-9
    assert abs(ref[:, 9].sum() - ang[ok, 2].sum()) < 1e-9                      # the AoD/AoA error word of SURVEY 8e
    for r in range(world):
        np.testing.assert_allclose(got[r], ref, rtol=1e-13)
    db = par.nmse_db_per_cell(ref)
    assert db.shape == (4,) and np.all(np.isfinite(db))
