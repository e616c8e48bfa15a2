"""PhaseLift GPU vs oracle on a few instances (diagnostics; tests/test_gpu_phaselift.py is the gate)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import twoace_b200 as tw
from oracle import phaselift as opl

def rel(a, b):
    return tw.harness.aligned_rel_err(a, b)

def case(name, A, y, its, reduce=1):
    tr = opl.TfocsTrace()
    t = time.time()
    ref = opl.my_phase_lift(y, A, opl.TfocsOpts(maxIts=its), tr)
    t_or = time.time() - t
    t = time.time()
    sig, info = tw.phaselift_batch([A], [y], tw.PlOpts.default(maxIts=its, reduce=reduce))
    t_g = time.time() - t
    print(f"{name}: oracle niter {tr.niter} nprox {tr.n_prox} nbt {tr.n_backtracks} rank {tr.rank} L {tr.L:.4g} [{tr.status}] {t_or:.1f}s | "
          f"gpu {info[0].tolist()} {t_g:.2f}s | rel err {rel(sig[0], ref):.3e}", flush=True)

rng = np.random.default_rng(5)
# small Gaussian, m > n (no reduction)
n, m = 16, 96
A = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(2)
x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
y = np.abs(A @ x) ** 2
for its in (1, 2, 5, 30, 250, 1000):
    case(f"gauss n16 m96 its{its}", A, y, its)
# m < n: reduction
n, m = 40, 24
A = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(2)
x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
y = np.abs(A @ x) ** 2
for its in (1, 3, 30, 300):
    case(f"gauss n40 m24 red its{its}", A, y, its, 1)
    case(f"gauss n40 m24 full its{its}", A, y, its, 0)
if len(sys.argv) > 1:
    cb = tw.harness.load_codebook('random_probe_cb_16x16')
    inst = tw.harness.make_instance(np.random.SeedSequence(58659179), cb, 128, 20.0)
    case("cb n256 m128 its60", inst.A, (inst.B / 2) ** 2, 60)
    case("cb n256 m128 its60 full", inst.A, (inst.B / 2) ** 2, 60, 0)
    case("cb n256 m128 full run", inst.A, (inst.B / 2) ** 2, 4000)
