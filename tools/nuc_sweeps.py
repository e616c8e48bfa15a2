"""Jacobi sweeps per SVT-active iteration in real config-1 flows (stage words of the nuclear variant)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import twoace_b200 as tw
hz = tw.harness
cb = hz.load_codebook()
ctx = tw.Context(0)
p = tw.Params.default().fixed_iters()
for M in (32, 64, 128, 256):
    insts = hz.make_batch(8, cb, M, 20.0)
    res = tw.solve_batch(tw.NUCLEAR, [i.A for i in insts], [i.B for i in insts], 16, 16, [i.train_idx[:1] for i in insts], p, ctx)
    sw = res.stage_words          # [nb, 5, 16]
    it = sw[:, :, 2]
    print("M", M, "iters per stage", it.mean(axis=0), "sweeps per stage", sw[:, :, 8].mean(axis=0),
          "kcycles/iter: eig", (sw[:, :4, 9].sum() / it[:, :4].sum() / 1e3).round(1), "argminz", (sw[:, :4, 13].sum() / it[:, :4].sum() / 1e3).round(1),
          "xupd", (sw[:, :4, 10].sum() / it[:, :4].sum() / 1e3).round(1), "loop", (sw[:, :4, 11].sum() / it[:, :4].sum() / 1e3).round(1))
