"""Wall time of one entry-point call (8-value M sweep, 16x16) on the GPU."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import twoace_b200 as tw
from twoace_b200 import entrypoints as ep, harness as hz
cb = hz.load_codebook()
rng = np.random.default_rng(3)
_, vecH, _, _ = hz.generate_channel(rng, 16, 16, 3)
amp = np.abs(cb @ vecH) / 16 * 3e-5
rss_dbm = 10 * np.log10(amp ** 2 * 1000)
ctx = tw.Context(0)
for name in ("A2only", "A2nuclear", "phaselift"):
    fn = getattr(ep, "channel_recovery_ADMM_v2_simulation_" + name)
    t = time.time()
    a, g, info = fn(16, 16, np.abs(cb), np.angle(cb), rss_dbm, 1, ctx=ctx, details=True)
    dt = time.time() - t
    H = a[:, 0] * np.exp(1j * g[:, 0])
    print(name, "M", list(info["M"]), "wall %.2f s" % dt, "NMSE dB per M:",
          [round(10 * np.log10(hz.nmse(H[i], vecH) + 1e-300), 1) for i in range(len(info["M"]))], flush=True)
