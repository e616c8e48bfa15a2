import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import twoace_b200 as tw
from twoace_b200 import harness as hz, solvers as sv
cb = hz.load_codebook()
ctx = tw.Context(0)
insts = hz.make_batch(33, cb, 256, 20.0)
p = tw.Params.default().fixed_iters()
for tens in (0, 1, 0, 1):
    ctx.set_option("tensor", tens)
    for rep in range(3):
        ctx.set_timing(True)
        t0 = time.time()
        res = sv.solve_batch(tw.NUCLEAR, [i.A for i in insts], [i.B for i in insts], 16, 16, [i.train_idx[:1] for i in insts], p, ctx)
        dt = time.time() - t0
        ms_k, nl = ctx.timing_collect()
        print(f"tensor={tens} rep {rep}: wall {dt*1e3:.1f} ms, stage kernels {ms_k:.1f} ms, launches so far {ctx.launch_count}", flush=True)
