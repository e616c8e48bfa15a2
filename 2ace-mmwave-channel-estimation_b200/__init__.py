"""twoace_b200 — B200-native batched solver for 2ACE's ADMM CSI-recovery hot path.

``harness``  synthetic instances + metrics (host, NumPy)
``lib``      ctypes binding of libtwoace.so (CUDA, sm_100a) — no CPU fallback
``solvers``  mirror of the reference's MATLAB solver signatures over the C ABI
"""
from . import entrypoints, harness, lib, parallel, solvers, twostage  # noqa: F401
from .lib import MINL2, NUCLEAR, V1, V2, V3, V4, V4_MULTI, Context, Params, PlOpts, SynthParams, TwoaceError  # noqa: F401
from .solvers import (ADMM_v2, ADMM_v2_nuclear, MyPhaseLift, inferLowRank, inferLowRank_Nuclear, inferMinL2,  # noqa: F401
                      inferLowRankV2, inferLowRankV3, inferLowRankV4,
                      inferLowRankV4_multi, phaselift_batch, phaselift_batch_codebook, solve_batch,
                      solve_batch_codebook)
from .entrypoints import (channel_recovery_ADMM_v2_simulation_A2nuclear,  # noqa: F401
                          channel_recovery_ADMM_v2_simulation_directional,
                          channel_recovery_ADMM_v2_simulation_A2only,
                          channel_recovery_ADMM_v2_simulation_multiresolution,
                          channel_recovery_ADMM_v2_simulation_phaselift)
