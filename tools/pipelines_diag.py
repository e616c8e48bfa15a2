"""One GPU, one vs several pipelines (twoace_create_multi with the same device listed k times): host-buffer throughput of
the config-1 batch.  usage: pipelines_diag.py [k ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv

ks = [int(a) for a in sys.argv[1:]] or [1, 2, 3]
cb = hz.load_codebook()
base = tw.Context(0); base.set_codebook(cb)
Ms, snrs = [32, 64, 128, 256], [0.0, 10.0, 20.0, 30.0]
m, snr, tid = [], [], []
ci = 0
for M in Ms:
    for s in snrs:
        for t in range(32):
            m.append(M); snr.append(s); tid.append((ci << 32) | t)
        ci += 1
# interleave the cells so that every contiguous slice holds the same mix of M
order = np.argsort(np.arange(len(m)) % 32, kind="stable")
m, snr, tid = np.array(m)[order], np.array(snr)[order], np.array(tid)[order]
sp = tw.SynthParams.default()
inst = sv.synth_batch(m, snr, 0, cb.shape[0], tid, sp, base)
p = tw.Params.default().fixed_iters()
for k in ks:
    ctx = tw.Context([0] * k) if k > 1 else base
    if k > 1:
        ctx.set_codebook(cb)
    for rep in range(3):
        t0 = time.perf_counter()
        res = sv.solve_batch_codebook(tw.NUCLEAR, inst["rows"], 1 / 16, inst["B"], 16, 16, inst["train_idx"], p, ctx)
        dt = time.perf_counter() - t0
        print(f"pipelines {k} rep {rep}: {len(m) / dt:7.1f} solves/s ({dt * 1e3:.0f} ms)", flush=True)
    if k > 1:
        ctx.close()
