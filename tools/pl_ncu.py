"""Short PhaseLift run for ncu (maxIts bounded so that one kernel launch lasts ~0.2 s)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import twoace_b200 as tw
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
its = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cb = tw.harness.load_codebook('random_probe_cb_16x16')
ctx = tw.Context(0)
ctx.set_codebook(cb)
batch = tw.harness.make_batch(nb, cb, 128, 20.0)
sig, info = tw.phaselift_batch_codebook([b.rows for b in batch], 1 / 16.0, [(b.B / 2.0) ** 2 for b in batch], 256,
                                        tw.PlOpts.default(maxIts=its), ctx)
print("prox", info[:, 1].sum(), "sweeps", info[:, 8].sum())
