"""Phase breakdown of the general kernel (admm_stage_kernel) on a config-5 shaped stage launch: 32x32 antennas
(n = 1024), synthetic 2-bit codebook, M = 190.  usage: gen_diag.py [nb] [iters] [r]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 148
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
r = int(sys.argv[3]) if len(sys.argv) > 3 else 20
TX = RX = 32
n = TX * RX
cb = hz._ROOTS[hz.random_beam_codes(np.random.default_rng(20231017), 4096, n)]
insts = hz.make_batch(nb, cb, 190, 20.0, Nt=TX, Nr=RX)
ctx = tw.Context(0)
rng = np.random.default_rng(0)
As, Bs, X0 = [], [], []
for i in insts:
    tr = i.train_idx[0]
    As.append(i.A[tr]); Bs.append(i.B[tr] / np.linalg.norm(i.B))
    X0.append((rng.standard_normal((n, r)) + 1j * rng.standard_normal((n, r))) / 32)
p = tw.Params.default(maxiter=iters).fixed_iters()
for rep in range(2):
    ctx.set_timing(True); ctx.timing_collect()
    _, _, _, W = sv.infer_admm_batch(As, Bs, X0, r > 1, False, TX, RX, p, ctx=ctx)
    ms, cnt = ctx.timing_collect()
    k = 1e3 * iters
    print("m %d r %d: %.1f ms (%d launches); per-iteration kcycles: x update %.0f  y update %.0f  ArgMinZ %.0f (eigensolver %.0f)"
          "  loop %.0f | set-up %.0f kcycles | sweeps/iter %.1f" %
          (As[0].shape[0], r, ms, cnt, W[:, 10].mean() / k, W[:, 12].mean() / k, W[:, 13].mean() / k, W[:, 9].mean() / k,
           W[:, 11].mean() / k, W[:, 14].mean() / 1e3, W[:, 8].mean() / iters))
