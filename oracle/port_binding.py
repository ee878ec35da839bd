"""ctypes binding to oracle/liboracle.so, the plain-C restatement (test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

CAND_DTYPE = np.dtype(
    {
        "names": ["freq", "snr", "drift", "sync", "shift", "m_type", "lin_drift", "V1", "V2", "p1", "p2"],
        "formats": ["<f4", "<f4", "<f4", "<f4", "<i4", "<i4", "<f4", "<f8", "<f8", "<i4", "<i4"],
        "offsets": [0, 4, 8, 12, 16, 20, 24, 24, 32, 40, 44],
        "itemsize": 48,
    }
)


class FdrParams(C.Structure):
    _fields_ = [
        ("fs", C.c_int), ("fl", C.c_int), ("spb", C.c_int), ("maxdrift", C.c_int), ("maxfreqs", C.c_int),
        ("halfbandwidth", C.c_int), ("cf", C.c_int), ("threshold", C.c_float),
        ("size", C.c_int), ("m", C.c_int), ("hpbm", C.c_int), ("n", C.c_int),
        ("df", C.c_float), ("min_snr", C.c_float), ("w", C.c_float * 4096),
    ]


class SdCall(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("lagmin", C.c_int), ("lagmax", C.c_int), ("lagstep", C.c_int),
        ("ifmin", C.c_int), ("ifmax", C.c_int), ("fstep", C.c_float),
        ("f1_in", C.c_float), ("shift_in", C.c_int), ("drift_in", C.c_float),
        ("f1_out", C.c_float), ("shift_out", C.c_int), ("sync_out", C.c_float),
        ("symbols", C.c_ubyte * 162), ("pad", C.c_ubyte * 2),
    ]


class FanoCall(C.Structure):
    _fields_ = [
        ("symbols", C.c_ubyte * 162), ("data", C.c_ubyte * 11), ("pad", C.c_ubyte * 3),
        ("result", C.c_int), ("metric", C.c_uint), ("cycles", C.c_uint), ("maxnp", C.c_uint),
    ]


class Trace(C.Structure):
    _fields_ = [
        ("calls", C.POINTER(SdCall)), ("max_calls", C.c_int), ("n_calls", C.c_int),
        ("fanos", C.POINTER(FanoCall)), ("max_fanos", C.c_int), ("n_fanos", C.c_int),
    ]


def build():
    """compile oracle/liboracle.so if missing or stale"""
    src = os.path.join(_HERE, "uwspr_oracle.c")
    if (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        L.orc_fdr_init.restype = C.c_int
        L.orc_fdr_init.argtypes = [vp] + [C.c_int] * 8
        L.orc_spectrogram.argtypes = [vp] * 4
        L.orc_power.argtypes = [vp] * 3
        L.orc_normalize_peaks.restype = C.c_int
        L.orc_normalize_peaks.argtypes = [vp] * 5
        L.orc_coarse.argtypes = [vp, vp, vp, C.c_int]
        L.orc_fdr_transform.restype = C.c_int
        L.orc_fdr_transform.argtypes = [vp] * 4
        L.orc_slm_frequency_drift.restype = C.c_float
        L.orc_slm_frequency_drift.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_slm_trajectory.restype = C.c_int
        L.orc_slm_trajectory.argtypes = [C.c_int, vp, vp, vp, vp]
        L.orc_sync_and_demodulate.argtypes = [
            vp, C.c_int, vp, vp, C.c_long, vp, vp, C.c_int, C.c_int, C.c_float, vp,
            C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int,
        ]
        L.orc_demodulate.restype = C.c_int
        L.orc_demodulate.argtypes = [C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_int]
        L.orc_demodulate_ex.restype = C.c_int
        L.orc_demodulate_ex.argtypes = [C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp]
        L.orc_deinterleave.argtypes = [vp]
        L.orc_interleave.argtypes = [vp]
        L.orc_encode.argtypes = [vp, vp, C.c_uint]
        L.orc_fano.restype = C.c_int
        L.orc_fano.argtypes = [vp, vp, vp, vp, vp, C.c_uint, C.c_int, C.c_uint]
        L.orc_channel_symbols.argtypes = [vp, vp]
        L.orc_sync_bit.restype = C.c_int
        L.orc_sync_bit.argtypes = [C.c_int]
        L.orc_sliding_window_count.restype = C.c_long
        L.orc_sliding_window_count.argtypes = [C.c_long, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_set_nonlinear_intended_t.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _iq(x):
    return np.ascontiguousarray(x, dtype=np.complex64)


class OracleFDR:
    """restatement of uwspr.FDR (lib/FDR_impl.cc)"""

    def __init__(self, fs=375, fl=45000, spb=256, maxdrift=0, maxfreqs=200, halfbandwidth=10, cf=1500, threshold=10):
        self.L = lib()
        self.p = FdrParams()
        rc = self.L.orc_fdr_init(C.byref(self.p), fs, fl, spb, maxdrift, maxfreqs, halfbandwidth, cf, threshold)
        if rc != 0:
            raise ValueError("parameters outside the valid domain of the reference")
        self.n, self.size, self.hpbm, self.m = self.p.n, self.p.size, self.p.hpbm, self.p.m
        self.df, self.min_snr = np.float32(self.p.df), np.float32(self.p.min_snr)
        self.fl, self.maxfreqs, self.cf = fl, maxfreqs, cf

    def window(self):
        return np.array(self.p.w[: self.size], dtype=np.float32)

    def spectrogram(self, x, want_spectra=False):
        x = _iq(x)
        assert x.size == self.fl
        ps = np.empty((self.n, self.size), np.float32)
        sp = np.empty((self.n, self.size), np.complex64) if want_spectra else None
        self.L.orc_spectrogram(C.byref(self.p), _p(x), _p(ps), _p(sp))
        return (ps, sp) if want_spectra else ps

    def power(self, spectra):
        spectra = np.ascontiguousarray(spectra, dtype=np.complex64)
        ps = np.empty((self.n, self.size), np.float32)
        self.L.orc_power(C.byref(self.p), _p(spectra), _p(ps))
        return ps

    def normalize_peaks(self, ps):
        ps = np.ascontiguousarray(ps, dtype=np.float32)
        psavg = np.empty(self.size, np.float32)
        smspec = np.empty(2 * self.hpbm, np.float32)
        cands = np.zeros(self.maxfreqs, CAND_DTYPE)
        npk = self.L.orc_normalize_peaks(C.byref(self.p), _p(ps), _p(psavg), _p(smspec), _p(cands))
        return cands[:npk].copy(), psavg, smspec

    def coarse(self, ps, cands):
        ps = np.ascontiguousarray(ps, dtype=np.float32)
        c = np.array(cands, dtype=CAND_DTYPE, copy=True)
        self.L.orc_coarse(C.byref(self.p), _p(ps), _p(c), len(c))
        return c

    def transform(self, x):
        x = _iq(x)
        cands = np.zeros(self.maxfreqs, CAND_DTYPE)
        npk = self.L.orc_fdr_transform(C.byref(self.p), _p(x), _p(cands), None)
        return cands[:npk].copy()


def sync_and_demodulate(cand, x, f1, shift1, drift1, mode, ifmin=0, ifmax=0, fstep=0.0, lagmin=0, lagmax=0,
                        lagstep=1, symfac=50, np_=45000, cf=1500):
    """restatement of sync_and_demodulate_impl::sync_and_demodulate; returns (f1, shift1, sync, symbols)"""
    L = lib()
    x = _iq(x)
    idat = np.ascontiguousarray(x.real)
    qdat = np.ascontiguousarray(x.imag)
    c = np.zeros(1, CAND_DTYPE)
    c[0] = cand
    symbols = np.zeros(162, np.uint8)
    f1c, sh, dr, sy = C.c_float(f1), C.c_int(shift1), C.c_float(drift1), C.c_float(0)
    L.orc_sync_and_demodulate(_p(c), cf, _p(idat), _p(qdat), np_, _p(symbols), C.byref(f1c), ifmin, ifmax,
                              C.c_float(fstep), C.byref(sh), lagmin, lagmax, lagstep, C.byref(dr), symfac,
                              C.byref(sy), mode)
    return np.float32(f1c.value), sh.value, np.float32(sy.value), symbols


def demodulate(x, cands, cf=1500, run_fano=True, max_calls=8192, max_fanos=4096):
    """restatement of sync_and_demodulate_impl::demodulate; returns (blobs, calls, fanos)"""
    L = lib()
    x = _iq(x)
    cands = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
    calls = (SdCall * max_calls)()
    fanos = (FanoCall * max_fanos)()
    tr = Trace(calls, max_calls, 0, fanos, max_fanos, 0)
    blobs = np.zeros((max(1, len(cands)), 7), np.uint8)
    nb = L.orc_demodulate(cf, _p(x), x.size, _p(cands), len(cands), C.byref(tr), _p(blobs), len(blobs), int(run_fano))
    assert tr.n_calls <= max_calls and tr.n_fanos <= max_fanos
    return blobs[:nb].copy(), [calls[i] for i in range(tr.n_calls)], [fanos[i] for i in range(tr.n_fanos)]


def demodulate_full(x, cands, cf=1500):
    """every mode-2 evaluation of every gated candidate (decoder skipped, so all 17 jiggled
    shifts run); returns (refined [npk,5] float32, list per candidate of the 17 mode-2 calls or [])"""
    L = lib()
    x = _iq(x)
    cands = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
    max_calls = 32 * max(1, len(cands))
    calls = (SdCall * max_calls)()
    tr = Trace(calls, max_calls, 0, None, 0, 0)
    refined = np.zeros((len(cands), 5), np.float32)
    L.orc_demodulate_ex(cf, _p(x), x.size, _p(cands), len(cands), C.byref(tr), None, 0, 0, _p(refined))
    assert tr.n_calls <= max_calls
    per, k = [], 0
    for j in range(len(cands)):
        k += 2 + (2 if cands[j]["m_type"] == 0 else 0)
        jigs = []
        if refined[j, 4] != 0:
            k += 2
            jigs = [calls[k + t] for t in range(17)]
            k += 17
        per.append(jigs)
    assert k == tr.n_calls
    return refined, per


def slm_frequency_drift(V1, V2, p1, p2, cf, t):
    return np.float32(lib().orc_slm_frequency_drift(V1, V2, p1, p2, cf, t))


def slm_trajectory(k):
    V1, V2, p1, p2 = C.c_double(), C.c_double(), C.c_int(), C.c_int()
    ok = lib().orc_slm_trajectory(k, C.byref(V1), C.byref(V2), C.byref(p1), C.byref(p2))
    return (V1.value, V2.value, p1.value, p2.value) if ok else None


def deinterleave(sym):
    s = np.array(sym, dtype=np.uint8, copy=True)
    lib().orc_deinterleave(_p(s))
    return s


def interleave(sym):
    s = np.array(sym, dtype=np.uint8, copy=True)
    lib().orc_interleave(_p(s))
    return s


def encode(data):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    sym = np.zeros(len(data) * 16, np.uint8)
    lib().orc_encode(_p(sym), _p(data), len(data))
    return sym


def fano(symbols, delta=60, maxcycles=10000, nbits=81):
    s = np.ascontiguousarray(symbols, dtype=np.uint8)
    data = np.zeros(11, np.uint8)
    metric, cycles, maxnp = C.c_uint(), C.c_uint(), C.c_uint()
    r = lib().orc_fano(C.byref(metric), C.byref(cycles), C.byref(maxnp), _p(data), _p(s), nbits, delta, maxcycles)
    return r, data, metric.value, cycles.value, maxnp.value


def channel_symbols(msg7):
    msg = np.ascontiguousarray(msg7, dtype=np.uint8)
    assert msg.size == 7
    out = np.zeros(162, np.uint8)
    lib().orc_channel_symbols(_p(msg), _p(out))
    return out


def sync_vector():
    return np.array([lib().orc_sync_bit(i) for i in range(162)], dtype=np.uint8)


def sliding_window_count(nitems, chunk, fs=375, fl=45000, shift=9):
    return lib().orc_sliding_window_count(nitems, chunk, fs, fl, shift)
