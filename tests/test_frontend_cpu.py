"""The front-end's composite filter against the block-by-block restatement of the flowgraph's GNU Radio
blocks (oracle/gr_frontend.py): host-side algebra only, no GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-uwspr_b200"))

from oracle import gr_frontend as gf  # noqa: E402
from uwspr_b200 import binding as ub  # noqa: E402


def test_tap_designs_equal_the_firdes_restatement():
    """the binding's vectorised float64 designs against the loop-for-loop restatement of firdes.cc (float32 taps,
    as GNU Radio stores them): same lengths, same values to float32 rounding"""
    pairs = [(ub.firdes_band_pass(1.0, 12000.0, 1490.0, 1510.0, 10.0), gf.band_pass(1.0, 12000.0, 1490.0, 1510.0, 10.0)),
             (ub.firdes_low_pass(1.0, 12000.0, 1510.0, 10.0), gf.low_pass(1.0, 12000.0, 1510.0, 10.0)),
             (ub.resampler_taps(1, 32), gf.design_filter(1, 32, 0.4)),
             (ub.resampler_taps(3, 2), gf.design_filter(3, 2, 0.4))]
    assert [len(p[1]) for p in pairs[:3]] == [2891, 2891, 1051]
    for mine, ref in pairs:
        assert len(mine) == len(ref) and ref.dtype == np.float32
        assert np.abs(mine - ref).max() <= 2e-7 * np.abs(ref).max()
    # the gains the designs promise: unity at DC, at the band centre, and `interp` at DC
    assert abs(pairs[1][1].astype(np.float64).sum() - 1.0) < 1e-5
    n = np.arange(2891) - 1445
    assert abs((pairs[0][1] * np.cos(2 * np.pi * 1500.0 / 12000.0 * n)).sum() - 1.0) < 1e-5
    assert abs(pairs[3][1].astype(np.float64).sum() - 3.0) < 1e-4


def test_composite_filter_equals_the_block_cascade():
    """y[m] = sum_j g[j] x[32 m - j] e^{-i w (32 m - j)} with the composite taps g of flowgraph_taps() equals the
    cascade band-pass -> translate + low-pass -> resample evaluated block by block"""
    rng = np.random.default_rng(3)
    n_in = 32 * 1500
    t = np.arange(n_in) / 12000.0
    audio = 0.3 * np.cos(2 * np.pi * (1502.2 * t + 0.3 * t * t)) + 0.5 * np.cos(2 * np.pi * 1800.0 * t) + 0.2 * rng.standard_normal(n_in)
    want = gf.flowgraph_frontend(audio)
    g = ub.flowgraph_taps().astype(np.complex128)
    bb = audio * np.exp(-2j * np.pi * 1500.0 / 12000.0 * np.arange(n_in))
    got = gf.fir_causal(bb, g)[::32]
    assert got.shape == want.shape == (1500,)
    scale = np.sqrt((np.abs(want[600:]) ** 2).mean())
    assert scale > 0.05                                      # the in-band tone comes through, the 1800 Hz one does not
    assert np.abs(got - want).max() < 2e-6 * scale           # float32 taps on both sides, float64 arithmetic
    # the 1800 Hz tone (amplitude 0.5) and the noise outside the 20 Hz pass band are gone: what is left is the chirp,
    # a real cosine of amplitude 0.3 -> 0.15 after the mix-down
    assert abs(np.abs(want[600:]).mean() - 0.15) < 0.015 and np.abs(want[600:]).max() < 0.2


def test_resampler_polyphase_indexing():
    """interp > 1: filter q sees taps[q::interp], the input advances once per `interp` of counter (3/2 here): equals
    zero-stuffing by 3, filtering, keeping every 2nd"""
    rng = np.random.default_rng(5)
    x = rng.standard_normal(200) + 1j * rng.standard_normal(200)
    taps = gf.design_filter(3, 2, 0.4)
    got = gf.rational_resampler_ccc(x, 3, 2, taps)
    up = np.zeros(600, np.complex128)
    up[::3] = x
    want = gf.fir_causal(up, taps.astype(np.float64))[::2]
    assert len(got) == 300 and np.abs(got - want).max() < 1e-12
    # and a common factor is divided out first (6/4 == 3/2), as the python wrapper does
    assert np.abs(gf.rational_resampler_ccc(x, 6, 4) - gf.rational_resampler_ccc(x, 3, 2)).max() == 0.0
