"""Oracle of the on-device instance synthesis (oracle/synth.py) on the CPU: the Philox4x32-10 generator against the
known-answer vectors of the Random123 distribution, and the defining properties of the generated instance."""
import math

import numpy as np

from oracle import synth


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors: "philox4x32 10"
    kat = [([0, 0, 0, 0], (0, 0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, (0xffffffff, 0xffffffff), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0),
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        got = synth.philox4x32_10(np.array(ctr, dtype=np.uint64), key)
        assert [int(x) for x in got] == want
    # vectorised == scalar
    ctr = np.array([[i, 3, 7, 0] for i in range(5)], dtype=np.uint64)
    both = synth.philox4x32_10(ctr, (11, 22))
    for i in range(5):
        assert np.array_equal(both[i], synth.philox4x32_10(ctr[i], (11, 22)))


def test_instance_properties(codebook):
    n = codebook.shape[1]
    ins = synth.synth_instance(codebook, 64, 20.0, 0, codebook.shape[0], trial=5, ntrain=3)
    # probes: distinct rows of the range; train draws: distinct ids, floor(0.95 m) each
    assert len(set(ins["rows"].tolist())) == 64 and ins["rows"].min() >= 0 and ins["rows"].max() < codebook.shape[0]
    assert ins["train_idx"].shape == (3, 60)
    for t in range(3):
        assert len(set(ins["train_idx"][t].tolist())) == 60 and ins["train_idx"][t].max() < 64
    assert not np.array_equal(ins["train_idx"][0], ins["train_idx"][1])
    # Eq. 23 channel: ||vecH||^2 = sum over path pairs, equals Nt Nr for orthogonal paths; rank <= L; angles in range
    H = ins["vecH"].reshape(16, 16, order="F")
    assert np.linalg.matrix_rank(H, tol=1e-9) <= 3
    assert np.all(np.abs(ins["aod"]) <= 47.5) and np.all(np.abs(ins["aoa"]) <= 47.5)
    # noiseless limit: B = |A vecH|
    hi = synth.synth_instance(codebook, 64, 300.0, 0, codebook.shape[0], trial=5)
    A = codebook[hi["rows"]] / math.sqrt(n)
    assert np.allclose(hi["B"], np.abs(A @ hi["vecH"]), rtol=0, atol=1e-12)
    # an instance depends on (seed, trial) only
    again = synth.synth_instance(codebook, 64, 20.0, 0, codebook.shape[0], trial=5, ntrain=3)
    assert np.array_equal(again["rows"], ins["rows"]) and np.array_equal(again["B"], ins["B"])
    other = synth.synth_instance(codebook, 64, 20.0, 0, codebook.shape[0], trial=6, ntrain=3)
    assert not np.array_equal(other["rows"], ins["rows"])
    # row range of a resolution stage is honoured
    st = synth.synth_instance(codebook, 36, 20.0, 1000, 1500, trial=1)
    assert st["rows"].min() >= 1000 and st["rows"].max() < 1500


def test_statistics_of_the_stream(codebook):
    """Noise power and uniformity of the probe selection over many trials."""
    Bn, first = [], []
    for t in range(200):
        ins = synth.synth_instance(codebook, 32, 0.0, 0, 256, trial=t)
        A = codebook[ins["rows"]] / 16.0
        # |y|^2 - |A h|^2 has mean = noise power (1 at 0 dB) for circular noise
        Bn.append(np.mean(ins["B"] ** 2 - np.abs(A @ ins["vecH"]) ** 2))
        first.append(int(ins["rows"][0]))
    assert abs(np.mean(Bn) - 1.0) < 0.1
    assert len(set(first)) > 120          # 200 draws from 256 rows: ~139 distinct expected
