"""PhaseLift throughput probe: nb codebook instances (M rows), device-timed through the context timing hooks."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import twoace_b200 as tw

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 148
M = int(sys.argv[2]) if len(sys.argv) > 2 else 128
its = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
cb = tw.harness.load_codebook('random_probe_cb_16x16')
ctx = tw.Context(0)
ctx.set_codebook(cb)
batch = tw.harness.make_batch(nb, cb, M, 20.0)
rows = [b.rows for b in batch]
ys = [(b.B / 2.0) ** 2 for b in batch]
o = tw.PlOpts.default(maxIts=its)
tw.phaselift_batch_codebook(rows[:2], 1 / 16.0, ys[:2], 256, tw.PlOpts.default(maxIts=5), ctx)
ctx.set_timing(True)
t = time.time()
sig, info = tw.phaselift_batch_codebook(rows, 1 / 16.0, ys, 256, o, ctx)
wall = time.time() - t
ms, cnt = ctx.timing_collect()
nprox = info[:, 1].sum()
print(f"nb {nb} M {M}: kernel {ms:.0f} ms wall {wall:.2f}s -> {nb / (ms / 1e3):.2f} solves/s; iters mean {info[:,0].mean():.0f} "
      f"min {info[:,0].min():.0f} max {info[:,0].max():.0f}; prox total {nprox:.0f} -> {nprox / (ms / 1e3):.0f} prox/s; "
      f"status {np.bincount(info[:,3].astype(int))}; rank mean {info[:,4].mean():.1f}")
cyc = info[:, 9:14].sum(axis=0)
print("cycle shares: grad %.3f warm %.3f jacobi %.3f z/Az %.3f x/tests %.3f; cycles per prox %.0f; sweeps per prox %.2f" % (
    *(cyc / cyc.sum()), cyc.sum() / nprox, info[:, 8].sum() / nprox))
print("subproblem solves: %.3f of Jacobi cycles" % (info[:, 14].sum() / info[:, 11].sum()))
