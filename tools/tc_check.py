"""Tensor-core (tcgen05 int8) A-products vs the FP64 SIMT products and the oracle: per-stage iterate parity and the
in-kernel cycle timers (prints, no asserts).  usage: tc_check.py [M ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv
from oracle import admm

cb = hz.load_codebook()
ctx = tw.Context(0)
tx = rx = 16
Ms = [int(a) for a in sys.argv[1:]] or [32, 64, 128, 256]

def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))

for M in Ms:
    ins = hz.make_batch(1, cb, M, 20.0)[0]
    A, B, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
    tr = ins.train_idx[0]
    At, Bt = A[tr], B[tr]
    Xs = admm.spectral_initialize(At, Bt, 20)
    for (sbr, nuc, cs) in [(True, False, 2), (True, False, 4), (False, False, 4), (True, True, 2), (True, True, 4)]:
        for iters in (1, 10, 100):
            p = tw.Params.default(maxiter=iters, tol_rel=0.0, tol_abs=0.0)
            zfn = admm.argmin_z_nuclear if nuc else admm.argmin_z
            snap = {iters: None}
            admm.infer_admm(At, Bt, Xs, sbr, False, tx, rx, 0.0, 1e-3, 1.03, 0.0, 0.0, iters, None, None, zfn, None, snap)
            s = snap[iters]
            out = {}
            for tens in (0, 1):
                ctx.set_option("tensor", tens); ctx.set_option("fast_cs", cs)
                t0 = ctx.tensor_launch_count
                Xg, Yg, Sg, W = sv.infer_admm_batch([At], [Bt], [Xs], sbr, False, tx, rx, p, nuclear=nuc, ctx=ctx)
                out[tens] = (Sg[0], W[0], ctx.tensor_launch_count - t0)
            S0, W0, _ = out[0]
            S1, W1, ntc = out[1]
            print(f"M={M:3d} m={At.shape[0]:3d} sbr={int(sbr)} nuc={int(nuc)} cs={cs} it={iters:3d} tc_launch={ntc}: "
                  f"TC-vs-oracle X {rel(S1['X'], s['X']):.1e} Y {rel(S1['Y'], s['Y']):.1e} Z {rel(S1['Z'], s['Z']):.1e} | "
                  f"SIMT-vs-oracle X {rel(S0['X'], s['X']):.1e} | TC-vs-SIMT X {rel(S1['X'], S0['X']):.1e} | "
                  f"cyc/it Xupd {W0[10]/iters:.0f}->{W1[10]/iters:.0f} total {W0[11]/iters:.0f}->{W1[11]/iters:.0f}", flush=True)
ctx.set_option("tensor", 1); ctx.set_option("fast_cs", 2)
