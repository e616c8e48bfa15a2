"""inferMinL2 on the GPU vs the oracle, stage by stage from identical start points (default tolerances).  Diagnostic."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv
from oracle import admm

def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))

cb = hz.load_codebook(); ctx = tw.Context(0)
M = 529
insts = hz.make_batch(4, cb, M, 20.0)
rng = np.random.default_rng(3)
p = tw.Params.default()
for ins in insts:
    tr = rng.permutation(M)[:math.ceil(M * 0.95)].astype(np.int32)
    rng.standard_normal(M)
    An, Bn, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
    At, Bt = An[tr], Bn[tr]
    X0 = admm.spectral_initialize_minl2(At, Bt, 20)
    Xg0 = sv.spectral_init_batch([At], [Bt], 20, ctx)[0]
    print("spectral projector err", rel(Xg0 @ Xg0.conj().T, X0 @ X0.conj().T))
    for iters in (10, 50, 500):
        snap = {iters: None}
        ta = admm.StageTrace()
        Xa, Ya, _ = admm.infer_admm_minl2(At, Bt, X0, True, 0.0, 1e-4, 1e-8, iters, ta, snap)
        pp = tw.Params.default(maxiter=iters)
        Xga, Yga, Sg, W = sv.infer_admm_batch([At], [Bt], [X0], True, False, 16, 16, pp, nuclear=2, ctx=ctx)
        print(f"  stage A maxiter {iters}: iters {ta.iters}/{int(W[0][2])} out X err {rel(Xga[0], Xa):.2e}", "state X", rel(Sg[0]['X'], snap[iters]['X']) if snap[iters] else None)
    G = Xa.conj().T @ Xa
    Xb0 = Xa @ np.linalg.eigh(0.5 * (G + G.conj().T))[1]
    for iters in (10, 50, 100, 500):
        snap = {iters: None}
        tb = admm.StageTrace()
        Xb, Yb, _ = admm.infer_admm_minl2(At, Bt, Xb0, False, 0.0, 1e-4, 1e-8, iters, tb, snap)
        pp = tw.Params.default(maxiter=iters)
        Xgb, Ygb, Sg, W = sv.infer_admm_batch([At], [Bt], [Xb0], False, False, 16, 16, pp, nuclear=2, ctx=ctx)
        st = snap[iters]
        percol = [rel(Sg[0]['X'][:, c], st['X'][:, c]) for c in range(20)] if st else None
        print(f"  stage B maxiter {iters}: iters {tb.iters}/{int(W[0][2])} col {tb.opt_col}/{int(W[0][4])} out X err {rel(Xgb[0], Xb):.2e}",
              "per-column state err", np.array2string(np.array(percol), precision=1) if percol else None)
