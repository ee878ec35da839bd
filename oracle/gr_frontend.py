"""TEST INFRASTRUCTURE ONLY (imported by tests/ alone; the product never touches this file).

Block-by-block float64 restatement of the stock GNU Radio blocks that sit ahead of the path in the
reference's flowgraphs (examples/WaveFilePlusNoiseDecode.grc):

    blocks_float_to_complex                                  :  x[n] + 0j
    freq_xlating_fft_filter_ccc_0     (:834-893)  centre 0,      taps = band-pass  1490..1510 Hz, width 10, Hamming (:322-383)
    freq_xlating_fft_filter_ccc_0_0   (:894-958)  centre 1500,   taps = low-pass   1510 Hz,       width 10, Hamming (:384-420)
    rational_resampler_xxx_0          (:1753-1810) interp 1, decim 32, default taps (fractional_bw 0.4, Kaiser beta 7)

PARITY UNPINNED against GNU Radio itself: GNU Radio is a third-party dependency of the reference
(CMakeLists.txt:115 find_package(Gnuradio "3.7.2")), absent from /root/reference and from this image, so no
output of the real blocks exists to compare with.  What is restated here is their published algorithm
(GNU Radio 3.7 sources):

    gr-filter/lib/firdes.cc                       compute_ntaps, window, low_pass, band_pass (taps are float32)
    gr-filter/python/filter/freq_xlating_fft_filter.py
                                                  rtaps[i] = taps[i] e^{+i w i}, fft_filter_ccc(decim, rtaps),
                                                  then rotator_cc(-decim w), w = 2 pi centre / samp_rate
    gr-filter/lib/fft_filter_ccc_impl.cc          overlap-save FIR: y[n] = sum_k rtaps[k] x[n-k], zero history
    gr-blocks/lib/rotator_cc_impl.cc              y[n] = x[n] * phase, phase *= e^{i inc}, phase starts at 1
    gr-filter/python/filter/rational_resampler.py design_filter(interp, decim, fractional_bw)
    gr-filter/lib/rational_resampler_base_XXX_impl.cc.t
                                                  polyphase: filter q holds taps[q::interp]; per output: emit
                                                  firs[ctr].filter(in), ctr += decim, while ctr >= interp: ctr -= interp, in++

Every block is evaluated on its own, in float64, in the order the flowgraph connects them; the library's
front-end collapses the cascade into one complex filter (uwspr_b200/binding.py flowgraph_taps) and the tests
compare the two.  GNU Radio computes in float32 (and its fft_filter through FFTW), so even the real blocks
would agree with this chain only to float32 rounding: the tests state their tolerance.
"""
import math

import numpy as np
from scipy.signal import fftconvolve

WIN_HAMMING, WIN_KAISER = "hamming", "kaiser"


def _izero(x):
    """firdes.cc Izero: power series of the modified Bessel function I0, terms until they drop under 1e-21 * sum"""
    s, u, n, half = 1.0, 1.0, 1, x / 2.0
    while True:
        t = half / n
        n += 1
        u *= t * t
        s += u
        if u < 1e-21 * s:
            return s


def max_attenuation(win, beta):
    return {WIN_HAMMING: 53.0, WIN_KAISER: beta / 0.1102 + 8.7}[win]


def compute_ntaps(fs, transition_width, win, beta):
    ntaps = int(max_attenuation(win, beta) * fs / (22.0 * transition_width))
    if ntaps % 2 == 0:
        ntaps += 1
    return ntaps


def window(win, ntaps, beta):
    w = np.empty(ntaps, np.float32)
    m = ntaps - 1
    if win == WIN_HAMMING:
        for n in range(ntaps):
            w[n] = 0.54 - 0.46 * math.cos((2 * math.pi * n) / m)
    else:
        ibeta = 1.0 / _izero(beta)
        inm1 = 1.0 / m
        for n in range(ntaps):
            t = 2 * n * inm1 - 1
            w[n] = _izero(beta * math.sqrt(max(0.0, 1.0 - t * t))) * ibeta
    return w


def low_pass(gain, fs, cutoff, transition_width, win=WIN_HAMMING, beta=6.76):
    ntaps = compute_ntaps(fs, transition_width, win, beta)
    taps = np.empty(ntaps, np.float32)
    w = window(win, ntaps, beta)
    M = (ntaps - 1) // 2
    fwT0 = 2 * math.pi * cutoff / fs
    for n in range(-M, M + 1):
        if n == 0:
            taps[n + M] = fwT0 / math.pi * w[n + M]
        else:
            taps[n + M] = math.sin(n * fwT0) / (n * math.pi) * w[n + M]
    fmax = float(taps[M])
    for n in range(1, M + 1):
        fmax += 2 * float(taps[n + M])
    return (taps * np.float32(gain / fmax)).astype(np.float32)


def band_pass(gain, fs, low, high, transition_width, win=WIN_HAMMING, beta=6.76):
    ntaps = compute_ntaps(fs, transition_width, win, beta)
    taps = np.empty(ntaps, np.float32)
    w = window(win, ntaps, beta)
    M = (ntaps - 1) // 2
    fwT0 = 2 * math.pi * low / fs
    fwT1 = 2 * math.pi * high / fs
    for n in range(-M, M + 1):
        if n == 0:
            taps[n + M] = (fwT1 - fwT0) / math.pi * w[n + M]
        else:
            taps[n + M] = (math.sin(n * fwT1) - math.sin(n * fwT0)) / (n * math.pi) * w[n + M]
    fmax = float(taps[M])
    for n in range(1, M + 1):
        fmax += 2 * float(taps[n + M]) * math.cos(n * (fwT0 + fwT1) * 0.5)
    return (taps * np.float32(gain / fmax)).astype(np.float32)


def fir_causal(x, taps):
    """y[n] = sum_k taps[k] x[n-k] for 0 <= n < len(x), x = 0 before the stream starts (fft_filter's zero tail,
    a FIR block's zero history)"""
    return fftconvolve(np.asarray(x, np.complex128), np.asarray(taps, np.complex128))[:len(x)]


def freq_xlating_fft_filter_ccc(x, decim, taps, center_freq, samp_rate):
    phase_inc = (2.0 * math.pi * center_freq) / samp_rate
    rtaps = np.asarray(taps, np.float64) * np.exp(1j * phase_inc * np.arange(len(taps)))
    y = fir_causal(x, rtaps)[::decim]
    return y * np.exp(-1j * decim * phase_inc * np.arange(len(y)))   # rotator_cc(-decim * phase_inc), phase(0) = 1


def design_filter(interp, decim, fractional_bw):
    if fractional_bw >= 0.5 or fractional_bw <= 0:
        raise ValueError("Invalid fractional_bandwidth, must be in (0, 0.5)")
    beta, halfband = 7.0, 0.5
    rate = float(interp) / float(decim)
    if rate >= 1.0:
        trans_width = halfband - fractional_bw
        mid = halfband - trans_width / 2.0
    else:
        trans_width = rate * (halfband - fractional_bw)
        mid = rate * halfband - trans_width / 2.0
    return low_pass(interp, interp, mid, trans_width, WIN_KAISER, beta)


def rational_resampler_ccc(x, interp, decim, taps=None, fractional_bw=None):
    g = math.gcd(interp, decim)
    if taps is None:
        taps = design_filter(interp // g, decim // g, 0.4 if fractional_bw is None else fractional_bw)
    interp, decim = interp // g, decim // g
    taps = np.asarray(taps, np.float64)
    per = -(-len(taps) // interp)
    padded = np.zeros(per * interp)
    padded[:len(taps)] = taps
    firs = [padded[q::interp] for q in range(interp)]
    x = np.asarray(x, np.complex128)
    xp = np.concatenate([np.zeros(per - 1, np.complex128), x])   # history = taps per filter
    out, ctr, pos = [], 0, 0
    while pos + per <= len(xp):
        seg = xp[pos:pos + per]
        out.append(np.dot(firs[ctr][::-1], seg))                 # fir_filter: sum_k taps[k] in[ntaps-1-k]
        ctr += decim
        while ctr >= interp:
            ctr -= interp
            pos += 1
    return np.array(out, np.complex128)


def flowgraph_frontend(audio, samp_rate=12000.0, center=1500.0, half_bandwidth=10.0, decim=32):
    """the flowgraph's cascade on one channel of real audio; returns the 375-sps complex stream (float64)"""
    x = np.asarray(audio, np.float64) + 0j
    bp = band_pass(1.0, samp_rate, center - half_bandwidth, center + half_bandwidth, 10.0, WIN_HAMMING, 6.76)
    lp = low_pass(1.0, samp_rate, center + half_bandwidth, 10.0, WIN_HAMMING, 6.76)
    y1 = freq_xlating_fft_filter_ccc(x, 1, bp, 0.0, samp_rate)
    y2 = freq_xlating_fft_filter_ccc(y1, 1, lp, center, samp_rate)
    return rational_resampler_ccc(y2, 1, decim)
