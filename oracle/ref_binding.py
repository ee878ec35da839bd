"""ctypes binding to oracle/_ref/libref_harness.so (test infrastructure).

The harness wraps the UNMODIFIED reference sources compiled against stubs (see
oracle/ref_harness.cc and oracle/Makefile).  It exists only where /root/reference was
available at build time; `available()` says whether it is loadable.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref_harness.so")
_lib = None

CAND_DTYPE = np.dtype(
    {
        # lib/candidate_t.h:27-50 -- 48 bytes, align 8
        "names": ["freq", "snr", "drift", "sync", "shift", "m_type", "lin_drift", "V1", "V2", "p1", "p2"],
        "formats": ["<f4", "<f4", "<f4", "<f4", "<i4", "<i4", "<f4", "<f8", "<f8", "<i4", "<i4"],
        "offsets": [0, 4, 8, 12, 16, 20, 24, 24, 32, 40, 44],
        "itemsize": 48,
    }
)


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


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB_PATH)
        L.ref_fdr_new.restype = C.c_void_p
        L.ref_fdr_new.argtypes = [C.c_int] * 8
        L.ref_fdr_free.argtypes = [C.c_void_p]
        L.ref_fdr_dims.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.ref_fdr_window.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_fdr_transform.restype = C.c_int
        L.ref_fdr_transform.argtypes = [C.c_void_p] * 6
        L.ref_sd_new.restype = C.c_void_p
        L.ref_sd_new.argtypes = [C.c_int] * 6 + [C.c_char_p]
        L.ref_sd_free.argtypes = [C.c_void_p]
        L.ref_sd_eval.argtypes = [
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
            C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
            C.c_int, C.c_void_p, C.c_int,
        ]
        L.ref_sd_demodulate.restype = C.c_int
        L.ref_sd_demodulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_pipeline.restype = C.c_int
        L.ref_pipeline.argtypes = [C.c_void_p] * 7 + [C.c_int]
        L.ref_fano_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_set_skip_fano.argtypes = [C.c_int]
        L.ref_slm_frequency_drift.restype = C.c_float
        L.ref_slm_frequency_drift.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_float, C.c_float]
        L.ref_slm_generate.restype = C.c_int
        L.ref_slm_generate.argtypes = [C.c_void_p, C.c_int]
        L.ref_fano_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
        L.ref_fano_mettab.argtypes = [C.c_void_p]
        L.ref_fano_decode.restype = C.c_int
        L.ref_fano_decode.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_uint]
        L.ref_deinterleave.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_pr3.argtypes = [C.c_void_p]
        L.ref_sw_new.restype = C.c_void_p
        L.ref_sw_new.argtypes = [C.c_int] * 4
        L.ref_sw_free.argtypes = [C.c_void_p]
        L.ref_sw_work.restype = C.c_int
        L.ref_sw_work.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_unpk.restype = C.c_int
        L.ref_unpk.argtypes = [C.c_void_p] * 4
        L.ref_nhash.restype = C.c_uint
        L.ref_nhash.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
        assert L.ref_candidate_size() == 48
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _iq(x):
    x = np.ascontiguousarray(x, dtype=np.complex64)
    return x


class RefFDR:
    """uwspr.FDR of the reference (lib/FDR_impl.cc), one window per call."""

    def __init__(self, fs=375, fl=45000, spb=256, maxdrift=0, maxfreqs=200, halfbandwidth=10, cf=1500, threshold=10):
        self.L = lib()
        self.h = self.L.ref_fdr_new(fs, fl, spb, maxdrift, maxfreqs, halfbandwidth, cf, threshold)
        self.fl, self.maxfreqs = fl, maxfreqs
        n, size, hpbm, m = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        df, min_snr = C.c_float(), C.c_float()
        self.L.ref_fdr_dims(self.h, C.byref(n), C.byref(size), C.byref(hpbm), C.byref(m), C.byref(df), C.byref(min_snr))
        self.n, self.size, self.hpbm, self.m = n.value, size.value, hpbm.value, m.value
        self.df, self.min_snr = np.float32(df.value), np.float32(min_snr.value)

    def window(self):
        w = np.empty(self.size, np.float32)
        self.L.ref_fdr_window(self.h, _p(w))
        return w

    def transform(self, x, want_ps=False, spectra=None):
        x = _iq(x)
        assert x.size == self.fl
        cands = np.zeros(self.maxfreqs, CAND_DTYPE)
        ps = np.empty((self.n, self.size), np.float32) if want_ps else None
        psavg = np.empty(self.size, np.float32) if want_ps else None
        if spectra is not None:
            spectra = np.ascontiguousarray(spectra, dtype=np.complex64)
            assert spectra.shape == (self.n, self.size)
        npk = self.L.ref_fdr_transform(self.h, _p(x), _p(cands), _p(ps), _p(psavg), _p(spectra))
        if want_ps:
            return cands[:npk].copy(), ps, psavg
        return cands[:npk].copy()

    def __del__(self):
        try:
            self.L.ref_fdr_free(self.h)
        except Exception:
            pass


class RefSD:
    """uwspr.sync_and_demodulate of the reference (lib/sync_and_demodulate_impl.cc)."""

    def __init__(self, fs=375, fl=45000, spb=256, maxdrift=0, maxfreqs=200, cf=1500, logdir="/tmp"):
        self.L = lib()
        self.h = self.L.ref_sd_new(fs, fl, spb, maxdrift, maxfreqs, cf, logdir.encode())
        self.fl = fl

    def eval(self, cand, x, f1, shift1, drift1, mode, ifmin=0, ifmax=0, fstep=0.0, lagmin=0, lagmax=0, lagstep=1, symfac=50, np_=45000):
        """one call of sync_and_demodulate(); returns (f1, shift1, sync, symbols)"""
        x = _iq(x)
        idat = np.ascontiguousarray(x.real)
        qdat = np.ascontiguousarray(x.imag)
        c = np.zeros(1, CAND_DTYPE)
        c[0] = cand
        symbols = np.zeros(162, np.uint8)
        f1c, sh, dr, sy = C.c_float(f1), C.c_int(shift1), C.c_float(drift1), C.c_float(0)
        self.L.ref_sd_eval(self.h, _p(c), _p(idat), _p(qdat), np_, _p(symbols), C.byref(f1c), ifmin, ifmax,
                           C.c_float(fstep), C.byref(sh), lagmin, lagmax, lagstep, C.byref(dr), symfac, C.byref(sy), mode)
        return np.float32(f1c.value), sh.value, np.float32(sy.value), symbols

    def demodulate(self, x, cands, max_calls=8192, max_fanos=4096):
        """returns (blobs[nb,7] uint8, calls list, fanos list)"""
        x = _iq(x)
        cands = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
        calls = (SdCall * max_calls)()
        fanos = (FanoCall * max_fanos)()
        tr = Trace(calls, max_calls, 0, fanos, max_fanos, 0)
        blobs = np.zeros((max(1, len(cands)), 7), np.uint8)
        nb = self.L.ref_sd_demodulate(self.h, _p(x), _p(cands), len(cands), C.byref(tr), _p(blobs), len(blobs))
        assert tr.n_calls <= max_calls and tr.n_fanos <= max_fanos
        return blobs[:nb].copy(), [calls[i] for i in range(tr.n_calls)], [fanos[i] for i in range(tr.n_fanos)]

    def deinterleave(self, sym):
        s = np.array(sym, dtype=np.uint8, copy=True)
        self.L.ref_deinterleave(self.h, _p(s))
        return s

    def __del__(self):
        try:
            self.L.ref_sd_free(self.h)
        except Exception:
            pass


def pipeline(fdr, sd, x, max_calls=8192, max_fanos=4096):
    """FDR -> sync_and_demodulate as wired in the flowgraphs; returns (cands, blobs, calls, fanos)"""
    L = lib()
    x = _iq(x)
    cands = np.zeros(fdr.maxfreqs, CAND_DTYPE)
    npk = C.c_int()
    calls = (SdCall * max_calls)()
    fanos = (FanoCall * max_fanos)()
    tr = Trace(calls, max_calls, 0, fanos, max_fanos, 0)
    blobs = np.zeros((fdr.maxfreqs, 7), np.uint8)
    nb = L.ref_pipeline(fdr.h, sd.h, _p(x), _p(cands), C.byref(npk), C.byref(tr), _p(blobs), len(blobs))
    return cands[: npk.value].copy(), blobs[:nb].copy(), [calls[i] for i in range(tr.n_calls)], [fanos[i] for i in range(tr.n_fanos)]


def fano_stats(reset=False):
    s, n = C.c_double(), C.c_long()
    lib().ref_fano_stats(C.byref(s), C.byref(n), int(reset))
    return s.value, n.value


def slm_frequency_drift(V1, V2, p1, p2, cf, t):
    return np.float32(lib().ref_slm_frequency_drift(V1, V2, p1, p2, cf, t))


def slm_generate():
    out = np.zeros((200, 4), np.float64)
    n = lib().ref_slm_generate(_p(out), 200)
    return out[:n]


def fano_encode(data):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    sym = np.zeros(len(data) * 16, np.uint8)
    lib().ref_fano_encode(_p(sym), _p(data), len(data))
    return sym


def fano_mettab():
    out = np.zeros((2, 256), np.int32)
    lib().ref_fano_mettab(_p(out))
    return out


def fano_decode(symbols, delta=60, maxcycles=10000):
    s = np.array(symbols, dtype=np.uint8, copy=True)
    data = np.zeros(11, np.uint8)
    metric, cycles, maxnp = C.c_uint(), C.c_uint(), C.c_uint()
    r = lib().ref_fano_decode(_p(data), _p(s), C.byref(metric), C.byref(cycles), C.byref(maxnp), delta, maxcycles)
    return r, data, metric.value, cycles.value, maxnp.value


def pr3():
    out = np.zeros(162, np.uint8)
    lib().ref_pr3(_p(out))
    return out


class RefUnpacker:
    """uwspr.WSPR_unpacker's unpk_ (lib/helpers.cc:494-590) with its 32768-entry callsign hash table"""

    def __init__(self):
        self.hashtab = np.zeros(32768 * 13, np.uint8)

    def unpack(self, message7):
        m = np.ascontiguousarray(message7, dtype=np.uint8)
        clp = np.zeros(64, np.uint8)
        cs = np.zeros(32, np.uint8)
        noprint = lib().ref_unpk(_p(m), _p(self.hashtab), _p(clp), _p(cs))
        return noprint, bytes(clp).split(b"\0")[0].decode("latin1")


def nhash(key, initval=146):
    k = np.frombuffer(key, np.uint8).copy()
    return int(lib().ref_nhash(_p(k), len(k), initval))
