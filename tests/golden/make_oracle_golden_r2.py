#!/usr/bin/env python
"""Regenerate tests/golden/oracle_golden_r2.npz: outputs of the NumPy oracle for the round-2 components (instance
synthesis, AoD/AoA error, inferMinL2 stage, large-M V4 stage, SVD reduction of the two-stage recovery).

NOT reference outputs (the reference is MATLAB: parity unpinned); they pin the oracle against drift and give the GPU
tests fixed vectors.  Only well-conditioned quantities are stored (few iterations: see DESIGN.md section 2 for why long
runs of the column-wise stages cannot be pinned).  Run from the repository root:
    python tests/golden/make_oracle_golden_r2.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

SYNTH_CASES = [(64, 20.0, 0, 3968, 7), (361, 10.0, 100, 3900, (3 << 32) | 5), (4, 30.0, 0, 1984, 11)]


def compute():
    import twoace_b200 as tw
    from oracle import admm, metrics as om, synth, twostage
    hz = tw.harness
    cb = hz.load_codebook()
    out = {}
    for k, (M, snr, lo, hi, tid) in enumerate(SYNTH_CASES):
        s = synth.synth_instance(cb, M, snr, lo, hi, tid, ntrain=3)
        for key in ("rows", "train_idx", "B", "vecH", "aod", "aoa"):
            out[f"synth{k}/{key}"] = np.asarray(s[key])
    s = synth.synth_instance(cb, 64, 20.0, 0, 3968, 7)
    noisy = s["vecH"] + 0.2 * np.exp(1j * np.arange(256))
    out["angles/exact"] = np.array(om.evaluation_angles(s["vecH"], s["aod"], s["aoa"], 16, 16))
    out["angles/noisy"] = np.array(om.evaluation_angles(noisy, s["aod"], s["aoa"], 16, 16))
    out["angles/x_noisy"] = noisy
    # stage iterates after a few iterations at m_train > n (well conditioned)
    ins = hz.make_batch(1, cb, 361, 20.0)[0]
    A, B, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
    tr = ins.train_idx[0]
    At, Bt = A[tr], B[tr]
    X0 = admm.spectral_initialize(At, Bt, 20)
    snap = {5: None}
    admm.infer_admm_minl2(At, Bt, X0, True, 0.0, 0.0, 0.0, 5, None, snap)
    out["minl2_M361_it5/X"], out["minl2_M361_it5/Y"] = snap[5]["X"], snap[5]["Y"]
    snap = {5: None}
    admm.infer_admm(At, Bt, X0, True, False, 16, 16, 0.0, 1e-3, 1.03, 0.0, 0.0, 5, None, None, admm.argmin_z, None, snap)
    out["v4_M361_it5/X"], out["v4_M361_it5/Z"] = snap[5]["X"], snap[5]["Z"]
    out["stage_M361/train_idx"] = np.asarray(tr)
    out["minl2_rank_M10"] = np.int64(admm.spectral_initialize_minl2(*admm._preprocess(ins.A[:10], ins.B[:10], 1e-8)[:2], 10).shape[1])
    P, C, mcs = twostage.svd_reduction(ins.A[:121] @ (np.eye(256)[:, ::2]), 3)
    out["twostage/mCS"] = np.int64(mcs)
    out["twostage/PC"] = P @ C
    return out


if __name__ == "__main__":
    blob = compute()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden_r2.npz"), **blob)
    for k, v in blob.items():
        print(k, np.shape(v))
