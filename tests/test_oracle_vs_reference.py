"""Pins the C restatement against the reference's own object code (oracle/_ref: the
unmodified reference sources compiled against stubs).  Runs only where oracle/_ref was
built (this container); everywhere else the committed golden vectors stand in."""
import numpy as np
import pytest

from oracle import port_binding as ob
from oracle import ref_binding as rb
from oracle import testdata as td

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built")]


def _same_calls(a, b):
    return len(a) == len(b) and all(bytes(x) == bytes(y) for x, y in zip(a, b))


def _same_fanos(a, b):
    # byte 172 = data[10], never written by the reference decoder (Fano.cc:243-247)
    return len(a) == len(b) and all(bytes(x)[:172] == bytes(y)[:172] and bytes(x)[176:] == bytes(y)[176:] for x, y in zip(a, b))


@pytest.mark.parametrize("maxdrift,hbw,thr", [(0, 10, 10), (4, 10, 10), (2, 40, 3), (1, 100, 1000000)])
def test_pipeline_bit_exact_on_synthetic(maxdrift, hbw, thr):
    rf = rb.RefFDR(maxdrift=maxdrift, halfbandwidth=hbw, threshold=thr)
    of = ob.OracleFDR(maxdrift=maxdrift, halfbandwidth=hbw, threshold=thr)
    sd = rb.RefSD(maxdrift=maxdrift)
    assert (rf.n, rf.size, rf.hpbm, rf.m, rf.df, rf.min_snr) == (of.n, of.size, of.hpbm, of.m, of.df, of.min_snr)
    assert np.array_equal(rf.window(), of.window())
    for w, snr in enumerate([-8.0, -19.0, -25.0, -28.5]):
        x, _ = td.synth_window(11, w, snr_db=snr)
        rc, rps, rpsavg = rf.transform(x, want_ps=True)
        ops = of.spectrogram(x)
        assert np.array_equal(ops, rps)
        c0, opsavg, _ = of.normalize_peaks(ops)
        assert np.array_equal(opsavg, rpsavg)
        oc = of.coarse(ops, c0)
        assert td.canon_cands(oc).tobytes() == td.canon_cands(rc).tobytes()
        rblobs, rcalls, rfanos = sd.demodulate(x, rc)
        oblobs, ocalls, ofanos = ob.demodulate(x, oc)
        assert rblobs.tobytes() == oblobs.tobytes()
        assert _same_calls(rcalls, ocalls) and _same_fanos(rfanos, ofanos)


def test_two_signals_and_noise_only():
    rf, of, sd = rb.RefFDR(halfbandwidth=40), ob.OracleFDR(halfbandwidth=40), rb.RefSD()
    a, _ = td.synth_window(3, 0, snr_db=-12.0, f0=-20.0)
    b, _ = td.synth_window(3, 1, snr_db=-15.0, f0=17.0)
    n, _ = td.synth_window(3, 2, snr_db=-60.0)
    for x in (a + b - n * 0, n, np.zeros(45000, np.complex64)):
        rc = rf.transform(x)
        oc = of.transform(x)
        assert td.canon_cands(oc).tobytes() == td.canon_cands(rc).tobytes()
        rblobs, rcalls, rfanos = sd.demodulate(x, rc)
        oblobs, ocalls, ofanos = ob.demodulate(x, oc)
        assert rblobs.tobytes() == oblobs.tobytes() and _same_calls(rcalls, ocalls) and _same_fanos(rfanos, ofanos)


def test_sync_and_demodulate_function_direct():
    """arbitrary arguments, both drift models, lags that run off both ends of the buffer"""
    sd = rb.RefSD()
    x, _ = td.synth_window(5, 0, snr_db=-15.0, f0=2.2, drift=1.7, start=500)
    rng = np.random.default_rng(5)
    cand = np.zeros(1, ob.CAND_DTYPE)
    for trial in range(12):
        nonlinear = trial % 3 == 2
        cand["m_type"] = 1 if nonlinear else 0
        if nonlinear:
            cand["V1"], cand["V2"], cand["p1"], cand["p2"] = ob.slm_trajectory(int(rng.integers(0, 125)))
        f1 = float(rng.uniform(-5, 5))
        shift = int(rng.integers(-400, 4000))
        drift = 0.0 if nonlinear else float(rng.choice([0.0, 1.0, -2.5, 0.5]))
        mode = trial % 3 if not nonlinear else int(rng.integers(0, 3))
        kw = dict(ifmin=-2, ifmax=2, fstep=0.25, lagmin=shift - 128, lagmax=shift + 128, lagstep=64)
        r = sd.eval(cand[0], x, f1, shift, drift, mode, **kw)
        o = ob.sync_and_demodulate(cand[0], x, f1, shift, drift, mode, **kw)
        assert r[0].tobytes() == o[0].tobytes() and r[1] == o[1] and r[2].tobytes() == o[2].tobytes()
        assert np.array_equal(r[3], o[3])


def test_slm_exhaustive():
    for k in range(125):
        V1, V2, p1, p2 = ob.slm_trajectory(k)
        for t in range(0, 120):
            a = rb.slm_frequency_drift(V1, V2, p1, p2, 1500.0, float(t))
            b = ob.slm_frequency_drift(V1, V2, p1, p2, 1500.0, float(t))
            assert a.tobytes() == b.tobytes()


def test_code_tables_and_decoder():
    assert np.array_equal(rb.pr3(), ob.sync_vector())
    rng = np.random.default_rng(9)
    for trial in range(20):
        if trial % 2:
            soft = rng.integers(0, 256, 162).astype(np.uint8)  # times out
        else:  # a decodable word with a few corrupted symbols
            data = np.zeros(11, np.uint8)
            data[:7] = td.message_bytes(rng)
            soft = np.where(ob.encode(data)[:162] == 1, 190, 66).astype(np.uint8)
            soft[rng.integers(0, 162, 12)] = 128
        r = rb.fano_decode(soft, maxcycles=200)
        o = ob.fano(soft, maxcycles=200)
        assert r[0] == o[0] and r[2:] == o[2:]
        if r[0] == 0:  # on a time-out the reference returns uninitialised heap for unvisited nodes
            assert np.array_equal(r[1][:10], o[1][:10])


def test_sliding_window_against_reference_block():
    sw = rb.lib().ref_sw_new(375, 45000, 9, 2)
    rng = np.random.default_rng(1)
    stream = (rng.standard_normal(45000 + 3 * 3375) + 1j * rng.standard_normal(45000 + 3 * 3375)).astype(np.complex64)
    out = np.zeros(45000, np.complex64)
    got, pos = [], 0
    while pos < len(stream):
        n = min(1125, len(stream) - pos)
        chunk = np.ascontiguousarray(stream[pos:pos + n])
        if rb.lib().ref_sw_work(sw, chunk.ctypes.data, n, out.ctypes.data):
            got.append(out.copy())
        pos += n
    rb.lib().ref_sw_free(sw)
    assert len(got) == ob.sliding_window_count(len(stream), 1125) == 4
    for k, w in enumerate(got):
        assert np.array_equal(w, stream[k * 3375:k * 3375 + 45000])
