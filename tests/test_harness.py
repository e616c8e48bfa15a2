import numpy as np

from twoace_b200 import harness as hz


def test_codebook_fixture_properties(codebook):
    cb = codebook
    assert cb.shape == (3968, 256)
    assert np.all(np.isin(cb, [1, 1j, -1, -1j]))        # 4th roots of unity
    # every row is kron(tx_row, rx_row): reshape is rank one (processsing_codebook_random.m:56)
    for i in (0, 17, 3967):
        s = np.linalg.svd(cb[i].reshape(16, 16), compute_uv=False)
        assert s[1] < 1e-12 * s[0]


def test_other_codebooks_load():
    assert hz.load_codebook_codes("random_probe_cb_16x16_multires").shape == (9920, 256)
    assert hz.load_codebook_codes("random_probe_cb_16x16_multires_actual").shape == (9920, 256)
    assert hz.load_codebook_codes("directional_codebook_16x16").shape == (32, 32, 256)


def test_instances_are_deterministic_and_shard_independent(codebook):
    a = hz.make_batch(6, codebook, 64, 20.0)
    b = hz.make_batch(3, codebook, 64, 20.0, first_trial=3)
    for i in range(3):
        np.testing.assert_array_equal(a[3 + i].rows, b[i].rows)
        np.testing.assert_array_equal(a[3 + i].B, b[i].B)
        np.testing.assert_array_equal(a[3 + i].train_idx, b[i].train_idx)
    ins = a[0]
    assert ins.A.shape == (64, 256) and ins.train_idx.shape == (3, 60)
    assert len(set(ins.rows.tolist())) == 64
    np.testing.assert_allclose(np.linalg.norm(ins.A, axis=1), 1.0)   # signal power 1
    assert abs(np.linalg.norm(ins.vecH) ** 2 / 256 - 1) < 0.5


def test_channel_is_rank_L():
    rng = np.random.default_rng(0)
    H, v, aod, aoa = hz.generate_channel(rng, 16, 16, 3)
    s = np.linalg.svd(H, compute_uv=False)
    assert s[3] < 1e-10 * s[0]
    np.testing.assert_array_equal(v, H.reshape(-1, order="F"))   # vecH = vec(H), H is Nr x Nt
    assert np.all(np.abs(aod) <= 47.5) and np.all(np.abs(aoa) <= 47.5)


def test_metrics():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    assert hz.nmse(3.1 * np.exp(0.7j) * x, x) < 1e-25          # scale/phase invariant
    assert hz.aligned_rel_err(np.exp(1.3j) * x, x) < 1e-14
    assert abs(hz.nmse(np.zeros(32) + 1e-30, x) - 1) < 1      # garbage -> ~0 dB
    assert hz.nmse_db([0.1, 0.1]) == -10.0


def test_load_codebook_mat_round_trip(tmp_path, codebook):
    """A codebook written the way the reference ships it (.mat v5, variable cb) loads back with exact phases."""
    from scipy.io import savemat
    from twoace_b200 import harness as hz
    noisy = codebook[:50] * (1 + 0j)
    noisy = np.abs(noisy) * np.exp(1j * np.angle(noisy))          # cos(pi/2) = 6e-17 residues, as in the shipped files
    path = str(tmp_path / "cb.mat")
    savemat(path, {"cb": noisy}, format="5")
    cb = hz.load_codebook_mat(path)
    assert cb.shape == (50, 256) and np.array_equal(cb, codebook[:50])
    savemat(path, {"cb": noisy.reshape(5, 10, 256)}, format="5")
    assert hz.load_codebook_mat(path).shape == (50, 256)
