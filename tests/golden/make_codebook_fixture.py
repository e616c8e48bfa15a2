"""Generate the packed 2-bit codebook fixtures from the reference's .mat files.

Run in the build container only (needs /root/reference):
    python tests/golden/make_codebook_fixture.py

Every entry of the shipped codebooks is a 4th root of unity exp(1j*k*pi/2), stored
in the .mat with ~3e-16 floating-point residue (cos(pi/2) = 6.1e-17).  The fixture stores the phase
code k in 2 bits (4 codes per byte, little-endian within the byte, row-major over [row, n]); the
loader rebuilds the EXACT values {1, 1j, -1, -1j}, which differ from the .mat values by <= 4e-16.
"""
import os
import sys

import numpy as np
import scipy.io as sio

REF = "/root/reference/codebook/codebook_mat"
OUT = os.path.join(os.path.dirname(__file__), "..", "..", "2ace-mmwave-channel-estimation_b200", "data")


def pack(cb):
    cb2 = cb.reshape(-1, cb.shape[-1])
    k = np.round(np.angle(cb2) / (np.pi / 2)).astype(np.int64) % 4
    exact = np.array([1, 1j, -1, -1j])[k]
    err = np.abs(exact - cb2).max()
    assert err < 1e-15, err
    k = k.astype(np.uint8).reshape(-1, 4)
    packed = k[:, 0] | (k[:, 1] << 2) | (k[:, 2] << 4) | (k[:, 3] << 6)
    return packed.astype(np.uint8), err


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ["random_probe_cb_16x16", "random_probe_cb_16x16_multires",
                 "random_probe_cb_16x16_multires_actual", "directional_codebook_16x16"]:
        cb = sio.loadmat(os.path.join(REF, name + ".mat"))["cb"]
        packed, err = pack(cb)
        np.savez(os.path.join(OUT, name + ".u2.npz"), codes=packed, shape=np.array(cb.shape, dtype=np.int64))
        print(name, cb.shape, "max |exact - mat| =", err, "bytes", packed.nbytes)


if __name__ == "__main__":
    sys.exit(main())
