import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv
from oracle import admm
cb = hz.load_codebook(); ctx = tw.Context(0)
ctx.set_option("fast_cs", int(sys.argv[1]) if len(sys.argv) > 1 else 4)
for M in (64, 128, 256):
    ins = hz.make_batch(1, cb, M, 20.0)[0]
    A, B, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
    tr = ins.train_idx[0]; At, Bt = A[tr], B[tr]
    X0 = admm.spectral_initialize(At, Bt, 20)
    for iters in (3, 40):
        p = tw.Params.default(maxiter=iters, tol_rel=0.0, tol_abs=0.0)
        snap = {iters: None}
        admm.infer_admm(At, Bt, X0, True, False, 16, 16, 0.0, 1e-3, 1.03, 0.0, 0.0, iters, None, None, admm.argmin_z_nuclear, None, snap)
        f0 = ctx.fast_launch_count
        Xg, Yg, Sg, W = sv.infer_admm_batch([At], [Bt], [X0], True, False, 16, 16, p, nuclear=True, ctx=ctx)
        s = snap[iters]
        print(f"M={M} m={len(tr)} it={iters}: fast={ctx.fast_launch_count-f0} sweeps={int(W[0][8])} |Zin|~{np.linalg.norm(s['X']+s['N']/s['mu']):.3g} tau={1/s['mu']:.3g} "
              f"X err {np.linalg.norm(Sg[0]['X']-s['X'])/np.linalg.norm(s['X']):.2e} |Z| gpu {np.linalg.norm(Sg[0]['Z']):.3g} oracle {np.linalg.norm(s['Z']):.3g}")
