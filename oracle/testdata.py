"""Test inputs for the hot path (test infrastructure): fixture loaders, the 12 kHz ->
375 sps front-end used to derive the committed golden windows, and the seeded
synthetic WSPR window generator of SURVEY.md section 8(d).
"""
import os
import struct
import wave

import numpy as np

from . import port_binding as ob

FS = 375
FL = 45000
SPB = 256
SEED_BASE = 20190222


def load_c2(path):
    """.c2 file: 14-byte name, int32 type, float64 MHz, 45000 x (f32 I, f32 Q).
    The reference conjugates on load (lib/c2file_source_impl.cc:80-92)."""
    raw = open(path, "rb").read()
    name, ntrmin, dfreq = struct.unpack("<14sid", raw[:26])
    x = np.frombuffer(raw[26:26 + 8 * FL], dtype="<f4").reshape(-1, 2)
    return (x[:, 0] - 1j * x[:, 1]).astype(np.complex64)


def read_wav(path):
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2
        n, ch, rate = w.getnframes(), w.getnchannels(), w.getframerate()
        a = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, ch)
    return a.astype(np.float64) / 32768.0, rate


def frontend(audio, rate=12000, carrier=1500.0, decim=32, ntaps=513, cutoff=150.0):
    """real audio at 12 kHz -> complex baseband at 375 sps: mix down by `carrier`,
    windowed-sinc low-pass, keep every 32nd sample (the stock GNU Radio blocks of
    examples/WaveFilePlusNoiseDecode.grc:834-958,1753-1810 do the equivalent)."""
    n = np.arange(len(audio))
    bb = audio * np.exp(-2j * np.pi * carrier * n / rate)
    k = np.arange(ntaps) - (ntaps - 1) / 2
    h = np.sinc(2 * cutoff / rate * k) * np.hamming(ntaps)
    h /= h.sum()
    y = np.convolve(bb, h, mode="full")[(ntaps - 1) // 2:][: len(audio)]
    return y[::decim].astype(np.complex64)


def message_bytes(rng):
    """50 random payload bits packed MSB-first into 7 bytes (last 6 bits zero)"""
    bits = rng.integers(0, 2, 50, dtype=np.uint8)
    b = np.zeros(56, np.uint8)
    b[:50] = bits
    return np.packbits(b)


def modulate(chan_syms, f0=0.0, drift=0.0, start=375, amp=1.0, fl=FL):
    """continuous-phase 4-FSK: tone (sym-1.5)*375/256 Hz, 256 samples per symbol,
    linear drift of +-drift/2 over the frame (fp = f0 + (drift/2)(i-81)/81)"""
    x = np.zeros(fl, np.complex128)
    df = FS / SPB
    i = np.repeat(np.arange(162), SPB)
    f = f0 + (np.asarray(chan_syms, dtype=np.float64)[i] - 1.5) * df + (drift / 2.0) * (i - 81) / 81.0
    phase = 2 * np.pi * np.cumsum(f) / FS
    n0, n1 = max(0, start), min(fl, start + 162 * SPB)
    x[n0:n1] = amp * np.exp(1j * phase[n0 - start:n1 - start])
    return x


def synth_window(stream, window, snr_db=None, f0=None, drift=None, start=None, maxdrift=3.0, fl=FL, seed=SEED_BASE):
    """one synthetic window keyed by (seed, stream, window); returns (x complex64, meta dict).
    Noise: complex AWGN with per-sample variance (375/2500)/10^(SNR/10) (2500 Hz SNR convention)."""
    rng = np.random.default_rng([seed, stream, window])
    msg = message_bytes(rng)
    syms = ob.channel_symbols(msg)
    f0 = float(rng.uniform(-6, 6)) if f0 is None else f0
    drift_d = float(rng.uniform(-maxdrift, maxdrift)) if drift is None else drift
    start_d = int(375 + rng.integers(0, 2561)) if start is None else start
    snr = float(rng.uniform(-30, 0)) if snr_db is None else snr_db
    x = modulate(syms, f0, drift_d, start_d, fl=fl)
    sigma2 = (375.0 / 2500.0) / 10 ** (snr / 10.0)
    noise = rng.standard_normal(fl) + 1j * rng.standard_normal(fl)
    x = x + np.sqrt(sigma2 / 2.0) * noise
    meta = dict(msg=msg, f0=f0, drift=drift_d, start=start_d, snr=snr)
    return x.astype(np.complex64), meta


def synth_batch(nwin, stream=0, first=0, **kw):
    xs = np.empty((nwin, kw.get("fl", FL)), np.complex64)
    metas = []
    for w in range(nwin):
        xs[w], m = synth_window(stream, first + w, **kw)
        metas.append(m)
    return xs, metas


def synth_array(nchan, window, whales, stream=5, snr_db=-18.0, whale_gain=1.0, fl=FL, seed=SEED_BASE):
    """BASELINE.json configs[4]: one transmitted frame seen by `nchan` hydrophones --
    per-channel delay U{0..64} samples and gain U(0.05, 0.2), independent AWGN at `snr_db` relative to a
    unit-amplitude frame (2500 Hz convention) scaled with the mean gain, plus `whales` (375-sps complex,
    looped, per-channel circular offset) at `whale_gain` (the x1 / x0.1 ratio of
    examples/WaveFilePlusNoiseDecode.grc:586,637).  Returns (x [nchan, fl] complex64, meta)."""
    rng = np.random.default_rng([seed, stream, window, 64])
    msg = message_bytes(rng)
    syms = ob.channel_symbols(msg)
    f0 = float(rng.uniform(-6, 6))
    start = int(375 + rng.integers(0, 2400))
    delays = rng.integers(0, 65, nchan)
    gains = rng.uniform(0.05, 0.2, nchan)
    offs = rng.integers(0, len(whales), nchan)
    sigma2 = (375.0 / 2500.0) / 10 ** (snr_db / 10.0) * 0.125 ** 2
    x = np.empty((nchan, fl), np.complex64)
    for c in range(nchan):
        sig = modulate(syms, f0, 0.0, start + int(delays[c]), amp=float(gains[c]), fl=fl)
        noise = (rng.standard_normal(fl) + 1j * rng.standard_normal(fl)) * np.sqrt(sigma2 / 2.0)
        wh = np.take(whales, (offs[c] + np.arange(fl)) % len(whales))
        x[c] = (sig + noise + whale_gain * wh).astype(np.complex64)
    return x, dict(msg=msg, f0=f0, start=start, delays=delays, gains=gains)


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_path(name):
    return os.path.join(GOLDEN_DIR, name)


def canon_cands(c):
    """canonical form of a candidate_t array for byte comparison: the `drift` field is never
    written by the reference and, for linear candidates, the union bytes past m_linear.drift
    are stale -- both are zeroed"""
    c = np.array(c, copy=True)
    raw = c.view(np.uint8).reshape(len(c), 48)
    raw[:, 8:12] = 0
    lin = c["m_type"] == 0
    raw[lin, 28:48] = 0
    return c
