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


def test_angle_error_oracle_properties():
    """Evaluation_Recovery.m:85-146 restated on the angular spectrum of an H-domain estimate."""
    from oracle import metrics as om
    Nt = Nr = 16
    kph, aod_v, aoa_v, (u0, u1), (v0, v1) = om.angle_grid(Nt, Nr, 64, 64, 95.0)
    # the searching area maps to a centred index range of the 64-point grid (sin(47.5 deg) = 0.737)
    assert (u0, u1) == (v0, v1) and u0 + u1 == 64 and 44 <= u1 - u0 + 1 <= 50
    # a single on-grid path is found exactly: all six errors vanish
    u, v = 40, 22
    aod = np.rad2deg(np.arcsin(aod_v[u] / kph))
    aoa = np.rad2deg(np.arcsin(aoa_v[v] / kph))
    at = np.exp(-1j * aod_v[u] * np.arange(Nt)) / 4
    ar = np.exp(-1j * aoa_v[v] * np.arange(Nr)) / 4
    H = np.outer(ar, at.conj()) * (0.3 - 0.8j)
    errs = om.evaluation_angles(H.reshape(-1, order="F"), [aod], [aoa], Nt, Nr)
    assert np.allclose(errs, 0.0, atol=1e-9)
    # off-grid path: the error is bounded by the grid spacing in degrees near that angle
    errs = om.evaluation_angles(H.reshape(-1, order="F"), [aod + 0.4], [aoa - 0.3], Nt, Nr)
    assert abs(errs[0] - 0.4) < 1e-9 and abs(errs[1] - 0.3) < 1e-9 and errs[3] < 1e-9
    # a global phase / scale of the estimate changes nothing; NaN estimates give NaN
    e1 = om.evaluation_angles(H.reshape(-1, order="F") * (2.0 * np.exp(0.7j)), [aod], [aoa], Nt, Nr)
    assert np.allclose(e1, 0.0, atol=1e-9)
    assert all(np.isnan(om.evaluation_angles(np.full(256, np.nan + 0j), [aod], [aoa], Nt, Nr)))
