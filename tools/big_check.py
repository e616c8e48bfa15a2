"""Chunked cluster kernel (big_stage.cuh, 256 < m <= 1024) vs the general kernel and the oracle: per-stage iterate
parity and cycle counters (prints, no asserts).  usage: big_check.py [M ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv
from oracle import admm

cb = hz.load_codebook("random_probe_cb_16x16_multires")
ctx = tw.Context(0)
tx = rx = 16
Ms = [int(a) for a in sys.argv[1:]] or [361, 529, 1024]

def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))

for M in Ms:
    ins = hz.make_batch(1, cb, M, 20.0, row_range=(5952, 9920))[0]
    A, B, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
    tr = ins.train_idx[0]
    At, Bt = A[tr], B[tr]
    Xs = admm.spectral_initialize(At, Bt, 20)
    for (sbr, tol) in [(True, 0.0), (False, 0.0), (True, 1e-4)]:
        for iters in (1, 10, 60):
            p = tw.Params.default(maxiter=iters, tol_rel=tol, tol_abs=0.0 if tol == 0 else 1e-8)
            snap = {iters: None}
            tro = admm.StageTrace()
            admm.infer_admm(At, Bt, Xs, sbr, False, tx, rx, 0.0, 1e-3, 1.03, tol, 0.0 if tol == 0 else 1e-8, iters, None, None,
                            admm.argmin_z, tro, snap)
            s = snap.get(iters) or snap.get(tro.iters)
            out = {}
            for tens in (0, 1):
                ctx.set_option("tensor", tens)
                t0 = ctx.tensor_launch_count
                Xg, Yg, Sg, W = sv.infer_admm_batch([At], [Bt], [Xs], sbr, False, tx, rx, p, ctx=ctx)
                out[tens] = (Sg[0], W[0], ctx.tensor_launch_count - t0, Xg[0])
            S0, W0, _, X0o = out[0]
            S1, W1, ntc, X1o = out[1]
            it1 = max(int(W1[2]), 1)
            sx = (lambda k: rel(S1[k], s[k])) if s is not None else (lambda k: float('nan'))
            print(f"M={M:4d} m={At.shape[0]:4d} sbr={int(sbr)} tol={tol:g} it={iters:3d} tc_launch={ntc}: big-vs-oracle X {sx('X'):.1e} "
                  f"Y {sx('Y'):.1e} Z {sx('Z'):.1e} M {sx('M'):.1e} | big-vs-general X {rel(S1['X'], S0['X']):.1e} out {rel(X1o, X0o):.1e} | "
                  f"iters {int(W0[2])}/{int(W1[2])}/{tro.iters} optit {int(W0[3])}/{int(W1[3])} mu {W0[0]:.6g}/{W1[0]:.6g} | "
                  f"cyc/it general {W0[11]/max(int(W0[2]),1):.0f} big {W1[11]/it1:.0f} (Xupd {W1[10]/it1:.0f} AX+YM {W1[12]/it1:.0f} Z {W1[13]/it1:.0f}) setup {W1[15]:.3g}",
                  flush=True)
ctx.set_option("tensor", 1)
