"""Multi-GPU execution of the ADMM path: trials shard across ranks, one statistics all-reduce.

The reference's only parallel construct is ``parfor`` over Monte-Carlo trials with a gather of the
per-trial ``Evaluation`` array (Numerical_Simulation/main_programs/Vs_M_par.m:145-196).  Here rank k
owns the contiguous trial slice [k*T/W, (k+1)*T/W); instances never communicate, so the data path has
no collective.  After the local solves one ``all_reduce(SUM)`` (NCCL over NVLink on GPUs, gloo in the
CPU tests) combines a small float64 statistics tensor [cells x metrics]; trial seeds are children of
one SeedSequence indexed by the global trial id, so results do not depend on the GPU count.
"""
from __future__ import annotations

import numpy as np

METRICS = ("count", "sum_mse", "n_rank_one", "n_rollback", "sum_quality", "sum_iters", "sum_gain_ana", "sum_gain_dig",
           "sum_proj_err", "sum_aoda_err")


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice of [0,total) owned by ``rank``."""
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def local_stats(cell_ids, n_cells: int, mse, info, metrics=None, angle_metrics=None) -> np.ndarray:
    """Accumulate per-cell sums for the instances of this rank.
    cell_ids: int [nb] (e.g. index into the SNR x M grid); mse: [nb]; info: [nb,16] (twoace.h words);
    metrics: optional [nb,4] from twoace_metrics_batch (MSE_H, gain_ana, gain_dig, proj_error), summed over the
    instances with a finite MSE_H (Evaluation_H.m:81-115; the reference gathers them per trial, Vs_M_par.m:190)."""
    s = np.zeros((n_cells, len(METRICS)), dtype=np.float64)
    cell_ids = np.asarray(cell_ids)
    mse = np.asarray(mse, dtype=np.float64)
    ok = np.isfinite(mse)
    np.add.at(s[:, 0], cell_ids[ok], 1.0)
    np.add.at(s[:, 1], cell_ids[ok], mse[ok])
    np.add.at(s[:, 2], cell_ids, info[:, 2])
    np.add.at(s[:, 3], cell_ids, info[:, 3])
    np.add.at(s[:, 4], cell_ids, np.nan_to_num(info[:, 0]))
    np.add.at(s[:, 5], cell_ids, info[:, 15])
    if metrics is not None:
        metrics = np.asarray(metrics, dtype=np.float64)
        for k in range(3):
            np.add.at(s[:, 6 + k], cell_ids[ok], np.nan_to_num(metrics[ok, 1 + k]))
    if angle_metrics is not None:      # AoDA_Err of twoace_angle_metrics_batch (Evaluation_Recovery.m:144-151), degrees
        a = np.asarray(angle_metrics, dtype=np.float64)[:, 2]
        np.add.at(s[:, 9], cell_ids[ok], np.nan_to_num(a[ok]))
    return s


def local_stats_device(cell_ids, n_cells: int, info, metrics, angle_metrics=None):
    """local_stats on the device with torch ops (no host round trip): cell_ids int64 [nb], info [nb,16],
    metrics [nb,4], angle_metrics [nb,6] or None: CUDA tensors -> [n_cells, len(METRICS)] float64 on the same device and
    stream."""
    import torch
    mse = metrics[:, 0]
    ok = torch.isfinite(mse)
    okf = ok.to(torch.float64)
    z = torch.zeros_like(mse)
    cols = [okf, torch.where(ok, mse, z), info[:, 2], info[:, 3], torch.nan_to_num(info[:, 0]), info[:, 15]]
    cols += [torch.where(ok, torch.nan_to_num(metrics[:, 1 + k]), z) for k in range(3)]
    cols.append(torch.where(ok, torch.nan_to_num(angle_metrics[:, 2]), z) if angle_metrics is not None else z)
    payload = torch.stack(cols, dim=1)
    s = torch.zeros((n_cells, len(METRICS)), dtype=torch.float64, device=mse.device)
    return s.index_add_(0, cell_ids, payload)


def all_reduce_stats(stats: np.ndarray, device=None) -> np.ndarray:
    """Sum the statistics tensor over all ranks (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    t = torch.from_numpy(np.ascontiguousarray(stats))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def nmse_db_per_cell(stats: np.ndarray) -> np.ndarray:
    """10*log10(mean MSE_H) per cell (Plot_result.m:154)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return 10 * np.log10(stats[:, 1] / stats[:, 0])
