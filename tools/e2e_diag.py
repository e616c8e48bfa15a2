"""Host-buffer vs device-buffer call of the same codebook-mode solve: wall time per call (diagnostic for the e2e leg)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv

variant = getattr(tw, sys.argv[1] if len(sys.argv) > 1 else "V4_MULTI")
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
M = int(sys.argv[3]) if len(sys.argv) > 3 else 64
T = 3 if variant == tw.V4_MULTI else 1
cb = hz.load_codebook(); ctx = tw.Context(0); ctx.set_codebook(cb)
sp = tw.SynthParams.default(ntrain=T)
out = sv.synth_batch([M] * nb, 20.0, 0, 3968, list(range(nb)), sp, ctx)
m = np.full(nb, M, np.int32)
rows = np.ascontiguousarray(np.concatenate(out["rows"])); B = np.ascontiguousarray(np.concatenate(out["B"]))
tr = np.ascontiguousarray(np.concatenate([t.reshape(-1) for t in out["train_idx"]]).astype(np.int32))
p = tw.Params.default().fixed_iters()
X = np.empty(nb * 256, np.complex128); Y = np.empty(nb * M, np.complex128); q = np.empty(nb)
dev = torch.device("cuda", 0)
B_d = torch.from_numpy(B).to(dev); X_d = torch.empty(nb * 512, dtype=torch.float64, device=dev)
Y_d = torch.empty(nb * M * 2, dtype=torch.float64, device=dev); q_d = torch.empty(nb, dtype=torch.float64, device=dev)
for rep in range(4):
    t0 = time.perf_counter()
    ctx.solve_batch_codebook_raw(variant, tw.lib.MEM_HOST, nb, 16, 16, m, rows, 1 / 16, B, tr, p, X, Y, q, None, None)
    t1 = time.perf_counter()
    ctx.solve_batch_codebook_raw(variant, tw.lib.MEM_DEVICE, nb, 16, 16, m, rows, 1 / 16, B_d.data_ptr(), tr, p, X_d.data_ptr(), Y_d.data_ptr(), q_d.data_ptr(), None, None)
    t2 = time.perf_counter()
    ctx.synchronize()
    t3 = time.perf_counter()
    print(f"rep {rep}: host-buffer call {1e3*(t1-t0):8.1f} ms | device-buffer call returns after {1e3*(t2-t1):8.1f} ms, done after {1e3*(t3-t1):8.1f} ms", flush=True)
