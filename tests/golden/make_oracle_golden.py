#!/usr/bin/env python
"""Regenerate tests/golden/oracle_golden.npz: outputs of the NumPy oracle on small seeded instances.

These are NOT reference outputs (the reference is MATLAB and could not be run: parity unpinned);
they pin the ORACLE against silent drift between rounds and give the GPU tests fixed vectors to compare with
(tests/test_golden.py).  Run from the repository root:  python tests/golden/make_oracle_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def cases():
    """name -> (kind, A, B, tx, rx, train_idx, maxiter).  Small enough for seconds of CPU."""
    import twoace_b200 as tw
    hz = tw.harness
    cb = hz.load_codebook()
    out = {}
    for name, M, kind, T, it in (("v4_M36", 36, "V4", 1, 60), ("multi_M36", 36, "V4_MULTI", 3, 40),
                                 ("nuclear_M64", 64, "NUCLEAR", 1, 60), ("v3_M36", 36, "V3", 1, 60)):
        ins = hz.make_batch(3, cb, M, 20.0)[2]
        out[name] = (kind, ins.A, ins.B, 16, 16, ins.train_idx[:T], it)
    rng = np.random.default_rng(2024)
    A = np.exp(1j * (np.pi / 2) * rng.integers(0, 4, (40, 64))) / 8
    _, vecH, _, _ = hz.generate_channel(rng, 8, 8, 3)
    B = np.abs(A @ vecH)
    out["v1_8x8"] = ("V1", A, B, 8, 8, rng.permutation(40)[:38][None, :].astype(np.int32), 60)
    ins = hz.make_batch(1, cb, 40, 20.0)[0]
    out["phaselift_M40"] = ("PHASELIFT", ins.A, (ins.B / 2) ** 2, 16, 16, np.zeros((0, 0), np.int32), 25)
    return out


def run_oracle(kind, A, B, tx, rx, tr, it):
    from oracle import admm, phaselift
    if kind == "PHASELIFT":
        return phaselift.my_phase_lift(B, A, phaselift.TfocsOpts(maxIts=it)), np.zeros(0, complex), np.nan
    fn = {"V4": admm.infer_low_rank_v4, "V4_MULTI": admm.infer_low_rank_v4_multi, "NUCLEAR": admm.infer_low_rank_nuclear,
          "V3": admm.infer_low_rank_v3, "V1": admm.infer_low_rank_v1}[kind]
    p = admm.Params(maxiter=it).fixed_iters()
    return fn(A, B, tx, rx, p, train_idx=tr if kind == "V4_MULTI" else tr[0])


if __name__ == "__main__":
    blob = {}
    for name, (kind, A, B, tx, rx, tr, it) in cases().items():
        X, Y, q = run_oracle(kind, A, B, tx, rx, tr, it)
        blob[name + "/X"], blob[name + "/Y"], blob[name + "/quality"] = X, Y, np.float64(q)
        print(name, kind, "quality", q, "|X|", np.linalg.norm(X))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **blob)
