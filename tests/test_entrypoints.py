"""Entry-point wrappers: integer bookkeeping on CPU (bit-exact), full call on GPU against the oracle."""
import numpy as np
import pytest

from twoace_b200 import entrypoints as ep


def test_measurement_counts_match_reference_list():
    # [4 36 121 225 361 529 784 1024] for 16x16 (A2only.m:106-118, main.py:67)
    np.testing.assert_array_equal(ep.measurement_counts(16, 16), [4, 36, 121, 225, 361, 529, 784, 1024])
    assert ep.measurement_counts(32, 32)[-1] == 4096 and ep.measurement_counts(4, 4)[0] == 4
    with pytest.raises(ValueError):
        ep.measurement_counts(5, 7)


def test_matlab_round_is_half_away_from_zero():
    np.testing.assert_array_equal(ep.matlab_round([0.5, 1.5, 2.5, -0.5, -2.5, 2.4]), [1, 2, 3, -1, -3, 2])


def test_multires_stage_rule():
    # …_multiresolution.m:137-143 with thresh = [96, 256], res_separation = [1984, 3968, 3968]
    assert ep.multires_row_range(4) == (0, 1984) and ep.multires_row_range(96) == (0, 1984)
    assert ep.multires_row_range(121) == (1984, 5952) and ep.multires_row_range(256) == (1984, 5952)
    assert ep.multires_row_range(361) == (5952, 9920) and ep.multires_row_range(1024) == (5952, 9920)


def test_rss_conversion():
    # sqrt(db2pow(-60 dBm)/1000) * 1e5/3
    np.testing.assert_allclose(ep.rss_dbm_to_amplitude([-60.0]), [np.sqrt(1e-6 / 1000) * 1e5 / 3], rtol=1e-15)


@pytest.mark.gpu
def test_a2only_and_multires_entry_points_against_oracle(gpu_ctx):
    """Synthetic RSS (dBm) from an Eq. 23 channel through the multires codebook; the wrapper output must equal
    the oracle run on the same rows / train splits (the reference-determined instances to 1e-4)."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz
    from oracle import admm
    cb = hz.load_codebook("random_probe_cb_16x16_multires")
    rng = np.random.default_rng(7)
    _, vecH, _, _ = hz.generate_channel(rng, 16, 16, 3)
    y = np.abs(cb @ vecH / 16) * (3 / 1e5)                       # amplitude so that B = |A x| after rss_fct
    rss_dbm = 10 * np.log10(y ** 2 * 1000)
    p = tw.Params.default(maxiter=60)
    amp, ang, info = tw.channel_recovery_ADMM_v2_simulation_multiresolution(
        16, 16, np.abs(cb), np.angle(cb), rss_dbm, 3, params=p, ctx=gpu_ctx, details=True)
    assert amp.shape == (8, 1, 256) and ang.shape == (8, 1, 256)
    np.testing.assert_array_equal(info["M"], [4, 36, 121, 225, 361, 529, 784, 1024])
    for M, rows in zip(info["M"], info["rows"]):
        lo, hi = ep.multires_row_range(int(M))
        assert len(rows) == M and rows.min() >= lo and rows.max() < hi and len(set(rows.tolist())) == M
    H = amp[:, 0, :] * np.exp(1j * ang[:, 0, :])
    po = admm.Params(maxiter=60)
    n_checked = 0
    for k in (1, 2, 3):                                          # M = 36, 121, 225
        rows, tr = info["rows"][k], info["train_idx"][k]
        B = ep.rss_dbm_to_amplitude(rss_dbm[rows])
        Xo, _, _ = admm.infer_low_rank_v4_multi(cb[rows], B, 16, 16, po, train_idx=list(tr))
        Xp, _, _ = admm.infer_low_rank_v4_multi(cb[rows], B * (1 + 1e-14 * rng.standard_normal(B.shape)), 16, 16, po,
                                                train_idx=list(tr))
        self_sens = hz.aligned_rel_err(Xp, Xo)
        err = hz.aligned_rel_err(H[k] * ep.RSS_FCT, Xo)
        if self_sens <= 1e-6:                                    # reference-determined (see test_gpu_parity.py)
            assert err < 1e-4
            n_checked += 1
        else:                                                    # noise-decided in the reference itself: 100x its own response
            assert err <= max(1e-4, 100 * self_sens)
    assert n_checked >= 1, "no reference-determined instance among M = 36 / 121 / 225"
    assert np.all(np.isfinite(H))                                 # NaN -> 0 (A2only.m:176)


@pytest.mark.gpu
def test_phaselift_entry_point_against_oracle(gpu_ctx):
    """channel_recovery_ADMM_v2_simulation_phaselift on a 4x4-antenna 2-bit random codebook (M sweep
    4 ... 64 probes): per-M result equals the oracle's MyPhaseLift with the scalings of Recover_Channel.m:35."""
    import twoace_b200 as tw
    from twoace_b200 import entrypoints as ep
    from twoace_b200 import harness as hz
    from oracle import phaselift as opl
    rng = np.random.default_rng(12)
    cb = np.exp(1j * (np.pi / 2) * rng.integers(0, 4, size=(200, 16)))
    h = (rng.standard_normal(16) + 1j * rng.standard_normal(16)) * 1e-5
    rss_dbm = 10 * np.log10(np.abs(cb @ h) ** 2 * 1000)
    o = tw.PlOpts.default(maxIts=60)
    amp, ang, info = ep.channel_recovery_ADMM_v2_simulation_phaselift(16 // 4, 4, np.abs(cb), np.angle(cb), rss_dbm, 1,
                                                                      opts=o, ctx=gpu_ctx, details=True)
    Ms = ep.measurement_counts(4, 4)
    assert amp.shape == (8, 1, 16) and list(info["M"]) == list(Ms) == [4, 9, 16, 25, 25, 36, 49, 64]
    for i, r in enumerate(info["rows"]):
        assert len(r) == Ms[i] and len(set(r.tolist())) == Ms[i]
        y = (ep.rss_dbm_to_amplitude(rss_dbm[r]) / 2e5) ** 2 * 1e10
        ref = opl.my_phase_lift(y, cb[r], opl.TfocsOpts(maxIts=60)) / np.sqrt(1e10) * 2e5 / ep.RSS_FCT
        got = amp[i, 0] * np.exp(1j * ang[i, 0])
        assert hz.aligned_rel_err(got, ref) < 1e-6
