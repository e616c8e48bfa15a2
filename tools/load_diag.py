"""Per-stage device-cycle breakdown of a full batch under load (all SMs busy), tensor-core vs FP64 SIMT products.
usage: load_diag.py [variant=NUCLEAR] [trials_per_M=32] [M ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv

variant = sys.argv[1] if len(sys.argv) > 1 else "NUCLEAR"
tpm = int(sys.argv[2]) if len(sys.argv) > 2 else 32
Ms = [int(a) for a in sys.argv[3:]] or [32, 64, 128, 256]
cb = hz.load_codebook()
ctx = tw.Context(0)
insts = []
for M in Ms:
    insts += hz.make_batch(tpm, cb, M, 20.0)
T = 3 if variant == "V4_MULTI" else 1
p = tw.Params.default().fixed_iters()
for tens in (0, 1):
    ctx.set_option("tensor", tens)
    for rep in range(2):
        ctx.set_timing(rep == 1)
        t0 = time.time()
        res = sv.solve_batch(getattr(tw, variant), [i.A for i in insts], [i.B for i in insts], 16, 16,
                             [i.train_idx[:T] for i in insts], p, ctx)
        dt = time.time() - t0
    ms_k, nl = ctx.timing_collect()
    print(f"tensor={tens}: stage kernels {ms_k:.1f} ms in {nl} launches")
    print(f"tensor={tens}: {len(insts)} solves in {dt:.2f} s (host buffers) tc launches {ctx.tensor_launch_count}")
    sw = np.asarray(res.stage_words)
    ms = np.array([len(i.B) for i in insts])
    for M in Ms:
        sel = ms == M
        w = sw[sel]                      # [inst, stage, word]
        for st in range(w.shape[1]):
            it = w[:, st, 2]
            act = it > 0
            if not act.any():
                continue
            tot = (w[act, st, 11] / it[act]).mean()
            print(f"   M={M:3d} stage {st}: {act.sum():3d} tasks, iters {it[act].mean():5.0f}, cycles/iter total {tot:8.0f} "
                  f"Xupd {(w[act, st, 10] / it[act]).mean():8.0f} YM {(w[act, st, 12] / it[act]).mean():7.0f} "
                  f"ArgMinZ {(w[act, st, 13] / it[act]).mean():8.0f} (eig {(w[act, st, 9] / it[act]).mean():8.0f}) "
                  f"tail {(w[act, st, 14] / it[act]).mean():7.0f} sweeps/iter {(w[act, st, 8] / it[act]).mean():.2f}")
