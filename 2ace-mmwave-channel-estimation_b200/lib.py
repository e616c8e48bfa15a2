"""ctypes binding of libtwoace (include/twoace.h).  There is no CPU fallback: if the CUDA shared
library is missing or no GPU is present the product path raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtwoace.so")

MEM_HOST, MEM_DEVICE = 0, 1
V4, V4_MULTI, NUCLEAR, V3, V2, V1, MINL2 = 0, 1, 2, 3, 4, 5, 6
INFO_WORDS = 16
STAGE_WORDS = 16
PL_INFO_WORDS = 16
METRIC_WORDS = 4
ANGLE_WORDS = 6

# every symbol include/twoace.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "twoace_default_params", "twoace_version", "twoace_create", "twoace_destroy", "twoace_last_error",
    "twoace_stream", "twoace_launch_count", "twoace_synchronize", "twoace_solve_batch",
    "twoace_set_codebook", "twoace_solve_batch_codebook", "twoace_infer_admm_batch",
    "twoace_spectral_init_batch", "twoace_set_timing", "twoace_timing_collect", "twoace_fp64_peak",
    "twoace_set_option", "twoace_fast_launch_count", "twoace_tensor_launch_count", "twoace_set_trace", "twoace_pl_default_opts", "twoace_phaselift_batch",
    "twoace_metrics_batch", "twoace_synth_default_params", "twoace_synth_batch",
    "twoace_create_multi", "twoace_device_count", "twoace_angle_metrics_batch",
]


class Params(C.Structure):
    """twoace_params (inferLowRankV4.m:2-9 defaults)."""
    _fields_ = [("lam", C.c_double), ("r", C.c_int32), ("mu0", C.c_double), ("rho", C.c_double),
                ("cc_frac", C.c_double), ("tol_rel", C.c_double), ("tol_abs", C.c_double),
                ("maxiter", C.c_int32)]

    @classmethod
    def default(cls, **kw) -> "Params":
        p = cls(0.0, 20, 1e-3, 1.03, 0.95, 1e-4, 1e-8, 500)
        for k, v in kw.items():
            if k == "lambda_":
                k = "lam"
            if not hasattr(p, k):
                raise TypeError(f"unknown solver parameter {k!r}")
            setattr(p, k, v)
        return p

    def fixed_iters(self) -> "Params":
        q = Params(self.lam, self.r, self.mu0, self.rho, self.cc_frac, 0.0, 0.0, self.maxiter)
        return q


class PlOpts(C.Structure):
    """twoace_pl_opts (MyPhaseLift.m:83-92 over the tfocs_initialize.m:9-40 defaults)."""
    _fields_ = [("maxIts", C.c_int32), ("tol", C.c_double), ("restart", C.c_int32), ("lam", C.c_double),
                ("alpha", C.c_double), ("beta", C.c_double), ("L0", C.c_double), ("cntr_reset", C.c_int32),
                ("backtrack_tol", C.c_double), ("reduce", C.c_int32)]

    @classmethod
    def default(cls, **kw) -> "PlOpts":
        o = cls(4000, 1e-10, 200, 5e-2, 0.9, 0.5, 1.0, 50, 1e-10, 1)
        for k, v in kw.items():
            if k == "lambda_":
                k = "lam"
            if not hasattr(o, k):
                raise TypeError(f"unknown PhaseLift option {k!r}")
            setattr(o, k, v)
        return o


class SynthParams(C.Structure):
    """twoace_synth_params: the numerical-simulation instance generator (Generate_Channel.m, Generate_Measurement.m)."""
    _fields_ = [("nt", C.c_int32), ("nr", C.c_int32), ("L", C.c_int32), ("searching_area", C.c_double),
                ("wavelength", C.c_double), ("spacing", C.c_double), ("row_scale", C.c_double),
                ("cc_frac", C.c_double), ("ntrain", C.c_int32), ("seed", C.c_uint64)]

    @classmethod
    def default(cls, nt: int = 16, nr: int = 16, **kw) -> "SynthParams":
        p = cls(nt, nr, 3, 95.0, 3e8 / 60.48e9, 3.055e-3, 1.0 / (nt * nr) ** 0.5, 0.95, 1, 58659179)
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown synthesis parameter {k!r}")
            setattr(p, k, v)
        return p


class TwoaceError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libtwoace.so (built in-tree by __graft_entry__.build()); fail loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TwoaceError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32p, dp = C.c_void_p, C.c_void_p, C.c_void_p
    lib.twoace_default_params.argtypes = [C.POINTER(Params)]
    lib.twoace_default_params.restype = None
    lib.twoace_version.restype = C.c_int
    lib.twoace_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.twoace_create.restype = C.c_int
    lib.twoace_create_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.twoace_create_multi.restype = C.c_int
    lib.twoace_device_count.argtypes = [vp]
    lib.twoace_device_count.restype = C.c_int
    lib.twoace_destroy.argtypes = [vp]
    lib.twoace_destroy.restype = None
    lib.twoace_last_error.argtypes = [vp]
    lib.twoace_last_error.restype = C.c_char_p
    lib.twoace_stream.argtypes = [vp]
    lib.twoace_stream.restype = vp
    lib.twoace_launch_count.argtypes = [vp]
    lib.twoace_launch_count.restype = C.c_int64
    lib.twoace_synchronize.argtypes = [vp]
    lib.twoace_synchronize.restype = C.c_int
    lib.twoace_solve_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, dp, dp, i32p,
                                       C.POINTER(Params), dp, dp, dp, dp, dp]
    lib.twoace_solve_batch.restype = C.c_int
    lib.twoace_set_codebook.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp]
    lib.twoace_set_codebook.restype = C.c_int
    lib.twoace_solve_batch_codebook.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p,
                                                C.c_double, dp, i32p, C.POINTER(Params), dp, dp, dp, dp, dp]
    lib.twoace_solve_batch_codebook.restype = C.c_int
    lib.twoace_infer_admm_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, i32p, dp, dp, C.c_int, dp,
                                            C.c_int, C.c_int, C.c_int, C.POINTER(Params), dp, dp, dp, dp]
    lib.twoace_infer_admm_batch.restype = C.c_int
    lib.twoace_spectral_init_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, i32p, dp, dp, C.c_int, dp]
    lib.twoace_spectral_init_batch.restype = C.c_int
    lib.twoace_set_timing.argtypes = [vp, C.c_int]
    lib.twoace_set_timing.restype = C.c_int
    lib.twoace_timing_collect.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.twoace_timing_collect.restype = C.c_int
    lib.twoace_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.twoace_fp64_peak.restype = C.c_int
    lib.twoace_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    lib.twoace_set_option.restype = C.c_int
    lib.twoace_fast_launch_count.argtypes = [vp]
    lib.twoace_fast_launch_count.restype = C.c_int64
    lib.twoace_set_trace.argtypes = [vp, C.c_int, vp, C.c_int64]
    lib.twoace_set_trace.restype = C.c_int
    lib.twoace_tensor_launch_count.argtypes = [vp]
    lib.twoace_tensor_launch_count.restype = C.c_int64
    lib.twoace_pl_default_opts.argtypes = [C.POINTER(PlOpts)]
    lib.twoace_pl_default_opts.restype = None
    lib.twoace_phaselift_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, i32p, dp, i32p, C.c_double, dp,
                                           C.POINTER(PlOpts), dp, dp]
    lib.twoace_phaselift_batch.restype = C.c_int
    lib.twoace_metrics_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp]
    lib.twoace_metrics_batch.restype = C.c_int
    lib.twoace_angle_metrics_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                               C.c_double, C.c_double, dp, dp, dp]
    lib.twoace_angle_metrics_batch.restype = C.c_int
    lib.twoace_synth_default_params.argtypes = [C.POINTER(SynthParams), C.c_int, C.c_int]
    lib.twoace_synth_default_params.restype = None
    lib.twoace_synth_batch.argtypes = [vp, C.c_int, C.c_int, C.POINTER(SynthParams), i32p, dp, i32p, i32p, vp, i32p, i32p,
                                       dp, dp, dp]
    lib.twoace_synth_batch.restype = C.c_int
    _lib = lib
    return lib


def _ptr(a):
    """Host pointer of a NumPy array, raw int for device pointers, None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


class Context:
    """twoace_ctx wrapper: one GPU (``Context(0)``) or several GPUs of a node behind one context
    (``Context([0, 1, 2, 3])``, twoace_create_multi: batches are split into per-GPU slices, host buffers only)."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            arr = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self.lib.twoace_create_multi(arr, len(device), C.byref(h))
        else:
            rc = self.lib.twoace_create(int(device), C.byref(h))
        if rc != 0 or not h.value:
            raise TwoaceError(f"twoace_create(device={device}) failed with code {rc}: no usable CUDA device "
                              "(the product path has no CPU fallback)")
        self.h = h
        self.device = device

    @property
    def device_count(self) -> int:
        return int(self.lib.twoace_device_count(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.twoace_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            msg = self.lib.twoace_last_error(self.h)
            raise TwoaceError(f"libtwoace error {rc}: {msg.decode() if msg else '?'}")

    @property
    def stream(self) -> int:
        return int(self.lib.twoace_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.twoace_launch_count(self.h))

    def synchronize(self):
        self.check(self.lib.twoace_synchronize(self.h))

    def set_option(self, key: str, value: int):
        self.check(self.lib.twoace_set_option(self.h, key.encode(), int(value)))

    @property
    def fast_launch_count(self) -> int:
        return int(self.lib.twoace_fast_launch_count(self.h))

    @property
    def tensor_launch_count(self) -> int:
        """Launches of the cluster kernel whose A-products ran on the tensor cores (tcgen05 kind::i8)."""
        return int(self.lib.twoace_tensor_launch_count(self.h))

    def set_trace(self, buf: "np.ndarray | None"):
        """Attach a host float64 array that the next solve_batch calls fill with res_comb per (instance, stage,
        iteration) -- see twoace_set_trace in include/twoace.h; None detaches it."""
        self._trace = buf
        if buf is None:
            self.check(self.lib.twoace_set_trace(self.h, MEM_HOST, None, 0))
        else:
            assert buf.dtype == np.float64 and buf.flags.c_contiguous
            self.check(self.lib.twoace_set_trace(self.h, MEM_HOST, buf.ctypes.data, buf.size))

    def set_timing(self, on: bool):
        self.check(self.lib.twoace_set_timing(self.h, int(bool(on))))

    def timing_collect(self):
        """(summed stage-kernel ms, stage-kernel launches) since the last collect."""
        ms, cnt = C.c_double(0.0), C.c_int64(0)
        self.check(self.lib.twoace_timing_collect(self.h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    def fp64_peak_tflops(self) -> float:
        v = C.c_double(0.0)
        self.check(self.lib.twoace_fp64_peak(self.h, C.byref(v)))
        return v.value

    # ---- raw entry points (pointers may be NumPy arrays (host) or ints (device addresses)) ----
    def solve_batch_raw(self, variant, mem, nb, tx, rx, m, A, B, train_idx, params, X, Y, quality, info=None,
                        stage_words=None):
        self.check(self.lib.twoace_solve_batch(self.h, variant, mem, nb, tx, rx, _ptr(m), _ptr(A), _ptr(B),
                                               _ptr(train_idx), C.byref(params), _ptr(X), _ptr(Y), _ptr(quality),
                                               _ptr(info), _ptr(stage_words)))

    def set_codebook(self, cb: np.ndarray):
        """cb: rows x n complex128 (any layout); uploaded column-major like the .mat variable."""
        cbf = np.asfortranarray(cb, dtype=np.complex128)
        flat = np.ascontiguousarray(cbf.reshape(-1, order="F"))
        self.check(self.lib.twoace_set_codebook(self.h, MEM_HOST, cb.shape[0], cb.shape[1], _ptr(flat)))

    def solve_batch_codebook_raw(self, variant, mem, nb, tx, rx, m, cb_rows, row_scale, B, train_idx, params, X, Y,
                                 quality, info=None, stage_words=None):
        self.check(self.lib.twoace_solve_batch_codebook(self.h, variant, mem, nb, tx, rx, _ptr(m), _ptr(cb_rows),
                                                        float(row_scale), _ptr(B), _ptr(train_idx),
                                                        C.byref(params), _ptr(X), _ptr(Y), _ptr(quality),
                                                        _ptr(info), _ptr(stage_words)))

    def infer_admm_batch_raw(self, mem, nb, tx, rx, m, A, B, r, X0, sbr, rank_one, nuclear, params, X, Y,
                             state=None, words=None):
        self.check(self.lib.twoace_infer_admm_batch(self.h, mem, nb, tx, rx, _ptr(m), _ptr(A), _ptr(B), r, _ptr(X0),
                                                    int(sbr), int(rank_one), int(nuclear), C.byref(params),
                                                    _ptr(X), _ptr(Y), _ptr(state), _ptr(words)))

    def phaselift_batch_raw(self, mem, nb, n, m, A, cb_rows, row_scale, y, opts, sig, info=None):
        self.check(self.lib.twoace_phaselift_batch(self.h, mem, nb, n, _ptr(m), _ptr(A), _ptr(cb_rows),
                                                   float(row_scale), _ptr(y), C.byref(opts), _ptr(sig), _ptr(info)))

    def metrics_batch_raw(self, mem, nb, tx, rx, X_est, X_true, phase_bit, out):
        self.check(self.lib.twoace_metrics_batch(self.h, mem, nb, tx, rx, _ptr(X_est), _ptr(X_true), int(phase_bit),
                                                 _ptr(out)))

    def angle_metrics_batch_raw(self, mem, nb, nt, nr, L, nqt, nqr, searching_area, wavelength, spacing, X_est, angles, out):
        self.check(self.lib.twoace_angle_metrics_batch(self.h, mem, nb, nt, nr, L, nqt, nqr, float(searching_area),
                                                       float(wavelength), float(spacing), _ptr(X_est), _ptr(angles), _ptr(out)))

    def synth_batch_raw(self, mem, nb, sp, m, snr_db, row_lo, row_hi, trial_id, cb_rows, train_idx, B, vecH, angles=None):
        self.check(self.lib.twoace_synth_batch(self.h, mem, nb, C.byref(sp), _ptr(m), _ptr(snr_db), _ptr(row_lo),
                                               _ptr(row_hi), _ptr(trial_id), _ptr(cb_rows), _ptr(train_idx), _ptr(B),
                                               _ptr(vecH), _ptr(angles)))

    def spectral_init_batch_raw(self, mem, nb, n, m, A, B, r, Xs):
        self.check(self.lib.twoace_spectral_init_batch(self.h, mem, nb, n, _ptr(m), _ptr(A), _ptr(B), r, _ptr(Xs)))


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
