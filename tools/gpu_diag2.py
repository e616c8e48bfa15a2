"""Full-solve parity diagnostic across M and variants (default tolerances)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv
from oracle import admm
cb = hz.load_codebook()
ctx = tw.Context(0)
ctx.set_option("fast", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
Ms = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 225]
names = sys.argv[3].split(",") if len(sys.argv) > 3 else ["V4", "NUCLEAR"]
snr = float(sys.argv[4]) if len(sys.argv) > 4 else 20.0
fixed = int(sys.argv[5]) if len(sys.argv) > 5 else 0
for M in Ms:
    insts = hz.make_batch(6, cb, M, snr)
    for name in names:
        variant = getattr(tw, name)
        T = 3 if name == "V4_MULTI" else 1
        p = tw.Params.default(); po = admm.Params()
        if fixed: p = tw.Params.default(maxiter=fixed).fixed_iters(); po = admm.Params(maxiter=fixed).fixed_iters()
        res = sv.solve_batch(variant, [i.A for i in insts], [i.B for i in insts], 16, 16, [i.train_idx[:T] for i in insts], p, ctx)
        fn = {"V4": admm.infer_low_rank_v4, "V4_MULTI": admm.infer_low_rank_v4_multi, "NUCLEAR": admm.infer_low_rank_nuclear}[name]
        for b, ins in enumerate(insts):
            info = admm.SolveInfo()
            tri = ins.train_idx[:3] if name == "V4_MULTI" else ins.train_idx[0]
            Xo, Yo, qo = fn(ins.A, ins.B, 16, 16, po, train_idx=tri, info=info)
            ran = [(int(w[2]), int(w[3]), int(w[4])) for w in res.stage_words[b] if w[2] > 0]
            oro = [(t.iters, t.opt_iter, t.opt_col) for t in info.traces]
            print(f"M={M} {name} inst {b}: err {hz.aligned_rel_err(res.X[b], Xo):.2e} q {res.quality[b]:.6f}/{qo:.6f} r1 {int(res.info[b,2])}/{int(info.used_rank_one)} "
                  f"rb {int(res.info[b,3])}/{int(info.rolled_back)} nmse {10*np.log10(hz.nmse(res.X[b], ins.vecH)):.2f}/{10*np.log10(hz.nmse(Xo, ins.vecH)):.2f}\n    gpu {ran}\n    ora {oro}")
