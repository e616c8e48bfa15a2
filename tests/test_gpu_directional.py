"""`directional` entry point (main/channel_recovery_ADMM_v2_simulation_directional.m) and the two-stage recovery behind
it (My_TwoStage_Recovery.m): stage I -- PhaseLift on the SVD-reduced mCS x mCS programme -- runs on the GPU and is compared
with the oracle (oracle/twostage.py -> oracle/phaselift.py); the host-side reduction and step II are checked by their
defining identities."""
import numpy as np
import pytest

from oracle import twostage as ots

pytestmark = pytest.mark.gpu


def _scene(seed=0):
    import twoace_b200 as tw
    from twoace_b200 import entrypoints as ep, harness as hz
    cb = hz.load_codebook("directional_codebook_16x16")                 # 32 x 32 x 256
    rng = np.random.default_rng(seed)
    _, vecH, _, _ = hz.generate_channel(rng, 16, 16, 3, searching_area=120.0, d=ep.DIRECTIONAL_SPACING)
    y = np.abs(cb.reshape(-1, 256) @ vecH / 16.0) * (1 + 0.01 * rng.standard_normal(1024))
    # dBm such that sqrt(db2pow(rss)/1000) * rss_fct == y   (…_directional.m:146)
    rss = 10 * np.log10(np.maximum(y / ep.RSS_FCT, 1e-12) ** 2 * 1000.0)
    return cb, rss.reshape(32, 32), vecH


@pytest.mark.parametrize("M_cur", [5, 11])
def test_two_stage_stage_one_matches_oracle(gpu_ctx, M_cur):
    import twoace_b200 as tw
    from twoace_b200 import entrypoints as ep, harness as hz, twostage as ts
    cb, rss, _ = _scene()
    idx = ep.directional_indexing(M_cur)
    beams = cb[np.ix_(idx, idx)].reshape(len(idx) ** 2, 256, order="F")
    meas = ep.rss_dbm_to_amplitude(rss[np.ix_(idx, idx)].reshape(-1, order="F"))
    yint = (meas / 2e5) ** 2 * 1e10
    AD = ep.sparse_dictionary(16, 16, 64, 64)
    A = beams @ AD
    plomp, plgamp, d = ts.my_two_stage_recovery(yint, A, 3, ctx=gpu_ctx, details=True)
    po, _, do = ots.my_two_stage_recovery(yint, A, 3)
    assert d["mCS"] == do["mCS"]                                                       # integer bookkeeping: bit-exact
    assert np.allclose(d["P"] @ d["C"], do["P"] @ do["C"], atol=1e-9)                  # same rank-mCS approximation
    # the leading eigenvector of the stage-I solution, up to its global phase.  TFOCS' stop is triggered by a 0/0 of
    # rounding residues in the reference itself ("PhaseLift"), so the two sides stop a few
    # iterations apart on these small, barely determined programmes: measured 1.5e-4 (M_cur = 5), bar 1e-3
    assert hz.aligned_rel_err(d["intSoln"], do["intSoln"]) < 1e-3
    # step II reproduces the stage-I vector in the reduced coordinates (OMP to a full support is an exact solve)
    assert np.linalg.norm(d["C"] @ plomp - d["intSoln"]) <= 1e-6 * np.linalg.norm(d["intSoln"])
    assert np.array_equal(plomp, plgamp)                                               # reference fallback branch


def test_directional_entry_point(gpu_ctx):
    import twoace_b200 as tw
    cb, rss, vecH = _scene(1)
    amp, ang, info = tw.channel_recovery_ADMM_v2_simulation_directional(16, 16, np.abs(cb), np.angle(cb), rss, 1,
                                                                         ctx=gpu_ctx, details=True,
                                                                         opts=tw.PlOpts.default(maxIts=400))
    assert amp.shape == (8, 2, 256) and ang.shape == (8, 2, 256)
    assert list(info["M"]) == [2, 6, 11, 15, 19, 23, 28, 32]
    assert [s["probes"] for s in info["stages"]] == [4, 36, 121, 225, 361, 529, 784, 1024]
    assert np.all(np.isfinite(amp)) and np.all(np.isfinite(ang))
    with pytest.raises(ValueError):
        tw.channel_recovery_ADMM_v2_simulation_directional(16, 16, np.abs(cb[:8]), np.angle(cb[:8]), rss, 1, ctx=gpu_ctx)
