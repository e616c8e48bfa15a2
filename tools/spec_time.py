"""Spectral initialisation: leading-eigenpair solver (tridiagonalisation + bisection + inverse iteration) vs the full
Jacobi decomposition, wall time of one batched call and agreement with numpy.linalg.eigh.  usage: spec_time.py [nb] [M ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv
from oracle import admm

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 148
Ms = [int(a) for a in sys.argv[2:]] or [121, 256, 361, 1024]
cb = hz.load_codebook()
ctx = tw.Context(0)
for M in Ms:
    insts = hz.make_batch(nb, cb, M, 20.0)
    As, Bs = [], []
    for i in insts:
        A, B, _, _ = admm._preprocess(i.A, i.B, 1e-8)
        tr = i.train_idx[0]
        As.append(A[tr]); Bs.append(B[tr])
    Xo = [admm.spectral_initialize(As[k], Bs[k], 20) for k in range(min(4, nb))]
    for jac in (1, 0):
        ctx.set_option("spectral_jacobi", jac)
        for rep in range(2):
            t0 = time.time()
            Xg = sv.spectral_init_batch(As, Bs, 20, ctx)
            dt = time.time() - t0
        err = max(np.linalg.norm(Xg[k] @ Xg[k].conj().T - Xo[k] @ Xo[k].conj().T) / np.linalg.norm(Xo[k] @ Xo[k].conj().T)
                  for k in range(len(Xo)))
        print(f"M={M:4d} m_train={As[0].shape[0]:4d} nb={nb} {'jacobi ' if jac else 'tridiag'}: {dt*1e3:8.1f} ms per call (host buffers), "
              f"max projector error vs eigh {err:.2e}", flush=True)
ctx.set_option("spectral_jacobi", 0)
