"""Stage-by-stage error amplification: GPU vs oracle on one instance (default tolerances)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv
from oracle import admm
cb = hz.load_codebook(); ctx = tw.Context(0)
ctx.set_option("fast", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
M = int(sys.argv[2]) if len(sys.argv) > 2 else 225
snr = float(sys.argv[3]) if len(sys.argv) > 3 else 20.0
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
def proj(X): return X @ X.conj().T
p = tw.Params.default()
for ins in hz.make_batch(3, cb, M, snr):
    A, B, An, Bn = admm._preprocess(ins.A, ins.B, 1e-8)
    tr = ins.train_idx[0]; At, Bt = A[tr], B[tr]
    Xs = admm.spectral_initialize(At, Bt, 20)
    Xg = sv.spectral_init_batch([At], [Bt], 20, ctx)[0]
    print(f"--- M={M}: spectral proj err {rel(proj(Xg), proj(Xs)):.1e}")
    for r1 in (False, True):
        # stage A from the ORACLE's Xs on both sides
        trA = admm.StageTrace()
        Xa, Ya, _ = admm.infer_admm(At, Bt, Xs, True, r1, 16, 16, 0.0, 1e-3, 1.03, 1e-4, 1e-8, 500, None, None, admm.argmin_z, trA)
        Xag, _, _, Wa = sv.infer_admm_batch([At], [Bt], [Xs], True, r1, 16, 16, p, ctx=ctx)
        # and from the GPU's own spectral init (gauge differs): compare gauge invariant X X'
        Xag2, _, _, Wa2 = sv.infer_admm_batch([At], [Bt], [Xg], True, r1, 16, 16, p, ctx=ctx)
        G = Xa.conj().T @ Xa; w, Vx = np.linalg.eigh(0.5 * (G + G.conj().T))
        print(f"  r1={int(r1)} stage A: same-X0 err {rel(Xag[0], Xa):.1e} iters {int(Wa[0][2])}/{trA.iters}; own-X0 proj err {rel(proj(Xag2[0]), proj(Xa)):.1e} iters {int(Wa2[0][2])}; "
              f"eig(X'X) ratio min/max {w[0]/w[-1]:.1e}, gaps min rel {np.min(np.diff(w))/w[-1]:.1e}")
        Xo = Xa @ Vx
        trB = admm.StageTrace()
        Xb, Yb, _ = admm.infer_admm(At, Bt, Xo, False, r1, 16, 16, 0.0, 1e-3, 1.03, 1e-4, 1e-8, 500, None, None, admm.argmin_z, trB)
        Xbg, _, _, Wb = sv.infer_admm_batch([At], [Bt], [Xo], False, r1, 16, 16, p, ctx=ctx)
        print(f"         stage B: same-X0 err {hz.aligned_rel_err(Xbg[0][:,0], Xb[:,0]):.1e} iters {int(Wb[0][2])}/{trB.iters} optit {int(Wb[0][3])}/{trB.opt_iter} col {int(Wb[0][4])}/{trB.opt_col}")
        # perturbation sensitivity of the ORACLE itself: stage B from Xo*(1+1e-14 noise)
        rng = np.random.default_rng(0)
        Xp = Xo * (1 + 1e-14 * rng.standard_normal(Xo.shape))
        trP = admm.StageTrace()
        Xbp, _, _ = admm.infer_admm(At, Bt, Xp, False, r1, 16, 16, 0.0, 1e-3, 1.03, 1e-4, 1e-8, 500, None, None, admm.argmin_z, trP)
        Xap, _, _ = admm.infer_admm(At, Bt, Xs * (1 + 1e-14 * rng.standard_normal(Xs.shape)), True, r1, 16, 16, 0.0, 1e-3, 1.03, 1e-4, 1e-8, 500, None, None, admm.argmin_z, None)
        print(f"         ORACLE self-sensitivity (1e-14 relative input noise): stage A out {rel(Xap, Xa):.1e}; stage B out {hz.aligned_rel_err(Xbp[:,0], Xb[:,0]):.1e} iters {trP.iters}/{trB.iters} col {trP.opt_col}/{trB.opt_col}")
