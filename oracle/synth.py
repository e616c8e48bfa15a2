"""CPU oracle of the on-device instance synthesis -- TEST INFRASTRUCTURE ONLY.

Restates, in NumPy, what csrc/synth.cuh builds per (trial, M, SNR) instance:

* Eq. 23 sparse multipath channel -- Numerical_Simulation/src/generate_channel/Generate_Channel.m:76-139
  (L > 1: no Rician tail, :98-106; vecH = vec(H_Matrix) with H_Matrix Nr x Nt, :139);
* probe selection without replacement from a codebook row range -- randperm of
  main/channel_recovery_ADMM_v2_simulation_A2only.m:137, resolution stage of ..._multiresolution.m:137-143;
* RSS amplitudes |FW vecH + noise|, signal power 1 -- generate_measurement/Generate_Measurement.m:84-101;
* the randsample draws of inferLowRankV4.m:36-37.

PARITY UNPINNED against MATLAB: its randn / randperm / randsample streams cannot be reproduced outside MATLAB
(SURVEY.md H1), so the stream is a documented counter-based generator instead, Philox4x32-10 (Salmon, Moraes, Dror,
Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), pinned here by the known-answer vectors of the
Random123 distribution (tests/test_oracle_synth.py).  Stream layout: see csrc/synth.cuh.
"""
from __future__ import annotations

import math

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: [..., 4] uint32 counters, key: (k0, k1).  Returns [..., 4] uint32."""
    c = np.asarray(ctr, dtype=np.uint64)
    c0, c1, c2, c3 = c[..., 0].copy(), c[..., 1].copy(), c[..., 2].copy(), c[..., 3].copy()
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _draw(count, stream, trial, seed):
    """Philox words of counters (index = 0..count-1, stream, trial lo, trial hi) under key = seed."""
    ctr = np.zeros((count, 4), dtype=np.uint64)
    ctr[:, 0] = np.arange(count, dtype=np.uint64)
    ctr[:, 1] = stream
    ctr[:, 2] = int(trial) & 0xFFFFFFFF
    ctr[:, 3] = (int(trial) >> 32) & 0xFFFFFFFF
    return philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def _u53(a, b):
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0


def _box_muller(w):
    u0, u1 = _u53(w[:, 0], w[:, 1]), _u53(w[:, 2], w[:, 3])
    rr = np.sqrt(-2.0 * np.log(1.0 - u0))
    return rr * np.cos(2.0 * np.pi * u1) + 1j * rr * np.sin(2.0 * np.pi * u1)


def _sample(count, pick, stream, trial, seed):
    """The `pick` smallest (32-bit key, index) pairs of `count` candidates, ascending: a uniform draw without
    replacement in random order (randperm(count, pick) / randsample(count, pick))."""
    w = _draw(count, stream, trial, seed)
    keys = (w[:, 0].astype(np.uint64) << np.uint64(32)) | np.arange(count, dtype=np.uint64)
    return (np.sort(keys)[:pick] & _MASK).astype(np.int64)


def synth_instance(cb, m, snr_db, row_lo, row_hi, trial, *, nt=16, nr=16, L=3, searching_area=95.0,
                   wavelength=3e8 / 60.48e9, spacing=3.055e-3, row_scale=None, cc_frac=0.95, ntrain=1,
                   seed=58659179):
    """One instance.  Returns dict(rows, train_idx [ntrain, floor(m cc_frac)], B, vecH, aod, aoa)."""
    n = nt * nr
    row_scale = 1.0 / math.sqrt(n) if row_scale is None else row_scale
    # Generate_Channel.m:76-106
    w = _draw(L, 0, trial, seed)
    aod = (_u53(w[:, 0], w[:, 1]) - 0.5) * searching_area
    aoa = (_u53(w[:, 2], w[:, 3]) - 0.5) * searching_area
    g = _box_muller(_draw(L, 1, trial, seed)) / math.sqrt(2.0)
    g = g / np.linalg.norm(g)
    # :108-139   H = sqrt(Nt Nr) ARx diag(g) ATx',  steering vectors exp(-j 2 pi / lambda d sin(theta) k) / sqrt(N)
    kph = 2.0 * np.pi / wavelength * spacing
    pt, pr = kph * np.sin(np.deg2rad(aod)), kph * np.sin(np.deg2rad(aoa))
    ATx = np.exp(-1j * pt[None, :] * np.arange(nt)[:, None]) / math.sqrt(nt)
    ARx = np.exp(-1j * pr[None, :] * np.arange(nr)[:, None]) / math.sqrt(nr)
    H = math.sqrt(nt * nr) * (ARx * g[None, :]) @ ATx.conj().T
    vecH = H.reshape(-1, order="F")
    # probes (A2only.m:137) and measurements (Generate_Measurement.m:84-101)
    rows = row_lo + _sample(row_hi - row_lo, m, 2, trial, seed)
    noise = math.sqrt(10.0 ** (-snr_db / 10.0) / 2.0) * _box_muller(_draw(m, 3, trial, seed))
    B = np.abs(row_scale * (np.asarray(cb)[rows, :] @ vecH) + noise)
    mtr = int(math.floor(m * cc_frac))
    train = np.stack([_sample(m, mtr, 4 + t, trial, seed) for t in range(ntrain)])
    return dict(rows=rows.astype(np.int32), train_idx=train.astype(np.int32), B=B, vecH=vecH, aod=aod, aoa=aoa)
