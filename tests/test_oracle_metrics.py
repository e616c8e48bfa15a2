"""CPU checks of oracle/metrics.py (Evaluation_H.m:81-115, Quantize_PS.m)."""
import numpy as np

from oracle import metrics as om


def _chan(seed, Nt=16, Nr=16):
    from twoace_b200 import harness as hz
    rng = np.random.default_rng(seed)
    _, vecH, _, _ = hz.generate_channel(rng, Nt, Nr, 3)
    return vecH, rng


def test_perfect_estimate():
    vecH, _ = _chan(0)
    mse, ga, gd, pe = om.evaluation_h(vecH * np.exp(0.7j) * 3.0, vecH, 16, 16)
    H = vecH.reshape(16, 16, order="F")
    s = np.linalg.svd(H, compute_uv=False)
    assert mse < 1e-28 and pe < 1e-13
    assert abs(gd - s[0]) < 1e-10 * s[0]            # unquantised leading singular vectors reach sigma_1
    assert 0.5 * s[0] < ga <= s[0] * (1 + 1e-12)    # 2-bit phases lose < 3 dB on a dominant path


def test_mse_matches_harness_definition_and_is_invariant():
    from twoace_b200 import harness as hz
    vecH, rng = _chan(1)
    x = vecH + 0.1 * (rng.standard_normal(256) + 1j * rng.standard_normal(256))
    m0 = om.evaluation_h(x, vecH, 16, 16)
    m1 = om.evaluation_h(x * (2.5 * np.exp(-1.1j)), vecH, 16, 16)
    assert abs(m0[0] - hz.nmse(x, vecH)) < 1e-15
    assert abs(m0[0] - m1[0]) < 1e-13 and abs(m0[2] - m1[2]) < 1e-10 and abs(m0[3] - m1[3]) < 1e-10


def test_quantize_ps_levels():
    q = om.quantize_ps(np.exp(1j * np.array([0.1, 1.5, 3.1, -3.1, -1.6, -0.8])), 2)
    lev = np.round(np.angle(q * np.sqrt(6)) / (np.pi / 2)).astype(int) % 4
    assert list(lev) == [0, 1, 2, 2, 3, 0] or list(lev) == [0, 1, 2, 2, 3, 3]
    assert np.allclose(np.abs(q), 1 / np.sqrt(6))
    # first-minimum rule at the wrap-around: angle = pi picks +pi, angles just above -pi pick -pi (same phasor)
    assert np.allclose(om.quantize_ps(np.array([-1.0 + 0j]), 2), [-1.0])


def test_nan_estimate_gives_nan():
    vecH, _ = _chan(2)
    x = vecH.copy()
    x[3] = np.nan
    assert all(np.isnan(v) for v in om.evaluation_h(x, vecH, 16, 16))
