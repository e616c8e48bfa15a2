"""First-contact GPU diagnostic: per-phase parity of the CUDA path against the oracle (prints, no asserts)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv
from oracle import admm

np.set_printoptions(linewidth=200, precision=4)
cb = hz.load_codebook()
ctx = tw.Context(0)
if len(sys.argv) > 1: ctx.set_option("fast", int(sys.argv[1]))
if len(sys.argv) > 2: ctx.set_option("fast_cs", int(sys.argv[2]))

def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))

insts = hz.make_batch(4, cb, 64, 20.0)
tx = rx = 16
# ---- 1. spectral init
for ins in insts[:2]:
    tr = ins.train_idx[0]
    A, B, An, Bn = admm._preprocess(ins.A, ins.B, 1e-8)
    Xo = admm.spectral_initialize(A[tr], B[tr], 20)
    Xg = sv.spectral_init_batch([A[tr]], [B[tr]], 20, ctx)[0]
    print("spectral: proj err", rel(Xg @ Xg.conj().T, Xo @ Xo.conj().T), "norms", np.linalg.norm(Xg), np.linalg.norm(Xo))

# ---- 2. stage parity at 1,2,10,100 iterations
ins = insts[0]
tr = ins.train_idx[0]
A, B, An, Bn = admm._preprocess(ins.A, ins.B, 1e-8)
At, Bt = A[tr], B[tr]
Xs = admm.spectral_initialize(At, Bt, 20)
CASES = [(True, Xs, False, False), (False, Xs, False, False), (True, Xs[:, :1], True, False),
         (True, Xs, False, True), (True, Xs, True, False), (False, Xs, False, True), (True, Xs[:, :1], False, True)]
if len(sys.argv) > 3 and sys.argv[3] == "nuc": CASES = [c for c in CASES if c[3]]
for (sbr, X0, r1, nuc) in CASES:
    for iters in (1, 2, 10, 100, 300, 500):
        p = tw.Params.default(maxiter=iters, tol_rel=0.0, tol_abs=0.0)
        snap = {iters: None}
        tro = admm.StageTrace()
        zfn = admm.argmin_z_nuclear if nuc else admm.argmin_z
        Xo, Yo, _ = admm.infer_admm(At, Bt, X0, sbr, r1, tx, rx, 0.0, 1e-3, 1.03, 0.0, 0.0, iters, None, None, zfn, tro, snap)
        Xg, Yg, Sg, W = sv.infer_admm_batch([At], [Bt], [X0], sbr, r1, tx, rx, p, nuclear=nuc, ctx=ctx)
        s = snap[iters]
        print(f"sbr={int(sbr)} r={X0.shape[1]} r1={int(r1)} nuc={int(nuc)} it={iters:4d}: "
              f"X {rel(Sg[0]['X'], s['X']):.2e} Z {rel(Sg[0]['Z'], s['Z']):.2e} N {rel(Sg[0]['N'], s['N']):.2e} "
              f"Y {rel(Sg[0]['Y'], s['Y']):.2e} M {rel(Sg[0]['M'], s['M']):.2e} | out X {rel(Xg[0], Xo):.2e} Y {rel(Yg[0], Yo):.2e} "
              f"| mu {W[0][0]:.6g}/{s['mu']:.6g} obj {W[0][1]:.6g}/{s['opt_obj']:.6g} optit {int(W[0][3])}/{tro.opt_iter} "
              f"col {int(W[0][4])}/{tro.opt_col} bumps {int(W[0][5])}/{tro.n_mu_bumps} sweeps {int(W[0][8])}")

if len(sys.argv) > 3 and sys.argv[3] == "nuc":
    print("fast launches:", ctx.fast_launch_count, "of", ctx.launch_count); sys.exit(0)
# ---- 3. convergence-test mode (default tolerances)
p = tw.Params.default()
tro = admm.StageTrace()
Xo, Yo, co = admm.infer_admm(At, Bt, Xs, True, False, tx, rx, 0.0, 1e-3, 1.03, 1e-4, 1e-8, 500, None, None, admm.argmin_z, tro)
Xg, Yg, Sg, W = sv.infer_admm_batch([At], [Bt], [Xs], True, False, tx, rx, p, ctx=ctx)
print("tol mode: iters", int(W[0][2]), tro.iters, "conv", int(W[0][6]), co, "X", rel(Xg[0], Xo))

# ---- 4. direct (non-Woodbury) branch: m > 0.62 n
insb = hz.make_batch(1, cb, 361, 20.0)[0]
A2, B2, _, _ = admm._preprocess(insb.A, insb.B, 1e-8)
tr2 = insb.train_idx[0]
Xs2 = admm.spectral_initialize(A2[tr2], B2[tr2], 20)
Xg2 = sv.spectral_init_batch([A2[tr2]], [B2[tr2]], 20, ctx)[0]
print("spectral m>n: proj err", rel(Xg2 @ Xg2.conj().T, Xs2 @ Xs2.conj().T))
for iters in (1, 10, 100):
    p = tw.Params.default(maxiter=iters, tol_rel=0.0, tol_abs=0.0)
    snap = {iters: None}
    Xo, Yo, _ = admm.infer_admm(A2[tr2], B2[tr2], Xs2, True, False, tx, rx, 0.0, 1e-3, 1.03, 0.0, 0.0, iters, None, None, admm.argmin_z, None, snap)
    Xg, Yg, Sg, W = sv.infer_admm_batch([A2[tr2]], [B2[tr2]], [Xs2], True, False, tx, rx, p, ctx=ctx)
    s = snap[iters]
    print(f"direct m={len(tr2)} it={iters}: X {rel(Sg[0]['X'], s['X']):.2e} Z {rel(Sg[0]['Z'], s['Z']):.2e} Y {rel(Sg[0]['Y'], s['Y']):.2e} out {rel(Xg[0], Xo):.2e}")

# ---- 5. full solves
for variant, name, fn in [(tw.V4, "V4", admm.infer_low_rank_v4), (tw.NUCLEAR, "NUC", admm.infer_low_rank_nuclear),
                          (tw.V4_MULTI, "MULTI", admm.infer_low_rank_v4_multi)]:
    p = tw.Params.default().fixed_iters()
    po = admm.Params().fixed_iters()
    T = 3 if variant == tw.V4_MULTI else 1
    t0 = time.time()
    res = sv.solve_batch(variant, [i.A for i in insts], [i.B for i in insts], tx, rx,
                         [i.train_idx[:T] for i in insts], p, ctx)
    tg = time.time() - t0
    for b, ins in enumerate(insts):
        info = admm.SolveInfo()
        tri = ins.train_idx[:3] if variant == tw.V4_MULTI else ins.train_idx[0]
        t0 = time.time()
        Xo, Yo, qo = fn(ins.A, ins.B, tx, rx, po, train_idx=tri, info=info)
        to = time.time() - t0
        print(f"{name} inst {b}: rel err {hz.aligned_rel_err(res.X[b], Xo):.2e} q {res.quality[b]:.6f}/{qo:.6f} "
              f"r1 {int(res.info[b,2])}/{int(info.used_rank_one)} rb {int(res.info[b,3])}/{int(info.rolled_back)} "
              f"best {int(res.info[b,4])}/{info.best_trial} nmse {10*np.log10(hz.nmse(res.X[b], ins.vecH)):.3f}/{10*np.log10(hz.nmse(Xo, ins.vecH)):.3f} dB "
              f"optit {[int(x) for x in res.stage_words[b,:,3]]} / {[t.opt_iter for t in info.traces]} oracle {to:.1f}s")
    print(f"{name}: gpu batch of {len(insts)} took {tg:.2f}s")

# ---- 6. throughput probe
for nbt in (148, 592):
    big = hz.make_batch(nbt, cb, 64, 20.0)
    p = tw.Params.default().fixed_iters()
    t0 = time.time()
    res = sv.solve_batch(tw.V4, [i.A for i in big], [i.B for i in big], tx, rx, [i.train_idx[:1] for i in big], p, ctx)
    dt = time.time() - t0
    print(f"throughput probe: {nbt} V4 solves in {dt:.2f}s = {nbt/dt:.1f} solves/s; total iters/inst {res.info[:,15].mean():.0f}; launches {ctx.launch_count}")
print("fast launches:", ctx.fast_launch_count, "of", ctx.launch_count)
