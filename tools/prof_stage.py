"""Small profiling driver: one batched InferADMM stage launch (for ncu) — prints the event-timed duration."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
r = int(sys.argv[3]) if len(sys.argv) > 3 else 20
sbr = int(sys.argv[4]) if len(sys.argv) > 4 else 1
M = int(sys.argv[5]) if len(sys.argv) > 5 else 64
cb = hz.load_codebook()
insts = hz.make_batch(nb, cb, M, 20.0)
ctx = tw.Context(0)
fast = int(sys.argv[6]) if len(sys.argv) > 6 else 1
cs = int(sys.argv[7]) if len(sys.argv) > 7 else 4
nuc = int(sys.argv[8]) if len(sys.argv) > 8 else 0
mu0 = float(sys.argv[9]) if len(sys.argv) > 9 else 1e-3
ctx.set_option("fast", fast); ctx.set_option("fast_cs", cs)
rng = np.random.default_rng(0)
As, Bs, X0 = [], [], []
for i in insts:
    tr = i.train_idx[0]
    As.append(i.A[tr]); Bs.append(i.B[tr] / np.linalg.norm(i.B))
    X0.append((rng.standard_normal((256, r)) + 1j * rng.standard_normal((256, r))) / 16)
p = tw.Params.default(maxiter=iters, mu0=mu0).fixed_iters()
for rep in range(2):
    ctx.set_timing(True); ctx.timing_collect()
    _, _, _, W = sv.infer_admm_batch(As, Bs, X0, bool(sbr), False, 16, 16, p, nuclear=bool(nuc), ctx=ctx)
    ms, cnt = ctx.timing_collect()
    if fast: print("  per-iteration kcycles: xupdate %.1f  yupdate %.1f  argminz %.1f (eig %.1f)  tail %.1f | loop total %.1f  sweeps/iter %.2f" % (W[:,10].mean()/iters/1e3, W[:,12].mean()/iters/1e3, W[:,13].mean()/iters/1e3, W[:,9].mean()/iters/1e3, W[:,14].mean()/iters/1e3, W[:,11].mean()/iters/1e3, W[:,8].mean()/iters))
    print(f"[fast={fast} cs={cs} fast_launches={ctx.fast_launch_count}] stage launch: nb={nb} iters={iters} r={r} sbr={sbr} M={M}: {ms:.2f} ms ({cnt} launches) -> {ms/iters*1e3:.1f} us/iter/batch, "
          f"{ms*1e-3/iters/ (nb/296.0) * 1.9e9/1e3:.0f} kcycles per iteration per CTA-slot wave")
