"""The C restatement (oracle/liboracle.so) against the golden vectors recorded from the
unmodified reference (tests/golden/make_golden.py).  CPU only; runs everywhere."""
import numpy as np
import pytest

from oracle import port_binding as ob
from oracle import testdata as td
from tests.conftest import case_window

CASES = ["ve3emb_c2", "test_1500", "rec_150613", "mix_whales", "syn_m0_w0", "syn_m0_w1", "syn_m0_w2", "syn_m0_w3",
         "syn_m0_w4", "syn_m0_w5", "syn_m4_w0", "syn_m4_w3", "syn_m4_w6", "syn_m4_w7"]


def _calls_bytes(calls):
    return [bytes(c) for c in calls]


def test_case_list_matches_golden(golden):
    assert list(golden["case_names"]) == CASES


@pytest.mark.parametrize("name", CASES)
def test_fdr_and_demodulate_bit_exact(name, golden, golden_windows):
    md = int(golden["case_maxdrift"][CASES.index(name)])
    x = case_window(name, golden_windows)
    f = ob.OracleFDR(maxdrift=md)
    ps = f.spectrogram(x)
    # spectrogram probes (the stub FFT of the reference build and the oracle's DFT are the same definition)
    assert np.array_equal(ps[[0, 1, 100, 347], :], golden[name + "/ps_rows"])
    assert ps.astype(np.float64).sum() == golden[name + "/ps_sum"][0]
    c0, psavg, _ = f.normalize_peaks(ps)
    assert np.array_equal(psavg, golden[name + "/psavg"])
    cands = f.coarse(ps, c0)
    want = golden[name + "/cands"]
    assert td.canon_cands(cands).view(np.uint8).reshape(-1, 48).tobytes() == want.tobytes()
    blobs, calls, fanos = ob.demodulate(x, cands)
    assert blobs.tobytes() == golden[name + "/blobs"].tobytes()
    gc = golden[name + "/calls"]
    assert len(calls) == len(gc)
    for a, b in zip(_calls_bytes(calls), gc):
        assert a == b.tobytes()
    gf = golden[name + "/fanos"]
    assert len(fanos) == len(gf)
    for a, b in zip(fanos, gf):
        assert bytes(a) == b.tobytes()


def test_fixture_messages(golden):
    # README.md:57-65 / SURVEY 8(c): VE3EMB FN25 30 and VE3EMB FN42 33
    assert golden["ve3emb_c2/blobs"].tobytes().hex() == "d42c73eb3a7780"
    assert golden["test_1500/blobs"].tobytes().hex() == "d42c73eb3a7780"
    assert golden["rec_150613/blobs"].tobytes().hex() == "d42c73eb0d1840"
    assert golden["mix_whales/blobs"].tobytes().hex() == "d42c73eb3a7780"


def test_slm_known_answers(golden):
    got = np.array([ob.slm_frequency_drift(1, -2.0, 0, 50, 1500.0, float(i)) for i in range(120)], np.float32)
    assert np.array_equal(got, golden["kat/slm_qa"])
    assert got[0] == 2.0 and got[20] == 0.0  # lib/slm_qa.cc sequence: 2, 1.97874, ..., 0 at t=20
    traj = np.array([ob.slm_trajectory(k) for k in range(125)])
    assert np.array_equal(traj, golden["kat/slm_traj"])
    assert ob.slm_trajectory(125) is None
    tab = np.array([[ob.slm_frequency_drift(t[0], t[1], int(t[2]), int(t[3]), 1500.0, float(k * 111 // 162))
                     for k in range(162)] for t in traj], np.float32)
    assert np.array_equal(tab, golden["kat/slm_table"])


def test_code_known_answers(golden):
    assert np.array_equal(ob.encode(golden["kat/encode_in"]), golden["kat/encode_out"])
    assert np.array_equal(ob.deinterleave(np.arange(162, dtype=np.uint8)), golden["kat/deinterleave_of_iota"])
    assert np.array_equal(ob.interleave(ob.deinterleave(np.arange(162, dtype=np.uint8))), np.arange(162))
    assert np.array_equal(ob.sync_vector(), golden["kat/pr3"])


def test_transmit_identity_on_c2_fixture(golden_windows):
    """tones of examples/VE3EMB.c2 == encode . interleave . (2*d + sync) of its decoded message"""
    x = golden_windows["ve3emb_c2"]
    chan = ob.channel_symbols(np.array([0xD4, 0x2C, 0x73, 0xEB, 0x3A, 0x77, 0x80], np.uint8))
    seg = x[375:375 + 162 * 256].reshape(162, 256)
    k = np.arange(256)
    tones = np.array([(s - 1.5) * 375.0 / 256.0 for s in range(4)])
    corr = np.abs(np.array([[np.sum(r * np.exp(-2j * np.pi * t * k / 375.0)) for t in tones] for r in seg]))
    assert np.array_equal(corr.argmax(axis=1), chan)


def test_fano_round_trip_and_timeout():
    rng = np.random.default_rng(7)
    for _ in range(5):
        msg = td.message_bytes(rng)
        data = np.zeros(11, np.uint8)
        data[:7] = msg
        enc = ob.encode(data)[:162]
        soft = np.where(enc == 1, 200, 56).astype(np.uint8)  # clean soft symbols
        r, dec, metric, cycles, maxnp = ob.fano(soft)
        assert r == 0 and np.array_equal(dec[:7], msg) and cycles == 82 and maxnp == 80
    r, dec, metric, cycles, maxnp = ob.fano(rng.integers(0, 256, 162).astype(np.uint8), maxcycles=100)
    assert r == -1 and cycles == 100 * 81 + 2


def test_domain_errors():
    with pytest.raises(ValueError):
        ob.OracleFDR(halfbandwidth=188)  # reference exits: lib/FDR_impl.cc:85-90
    with pytest.raises(ValueError):
        ob.OracleFDR(halfbandwidth=187)  # hazard H3: reads psavg[-3]


def test_sliding_window_count():
    # window k starts at k*shift*fs; one PDU per work() call at most (sliding_window...cc:113-135)
    assert ob.sliding_window_count(45000, 45000) == 1
    assert ob.sliding_window_count(45000 + 3375, 1125) == 2
    assert ob.sliding_window_count(44999, 4096) == 0


def test_array_fixture_decodes_on_the_oracle():
    """the hydrophone-array generator (BASELINE.json configs[4]) and its whale fixture: every channel
    of a small array carries the same decodable frame"""
    whales = np.load(td.golden_path("whales_375sps.npy"))
    assert whales.dtype == np.complex64 and whales.shape == (24241,)
    x, meta = td.synth_array(3, 0, whales, snr_db=-15.0)
    of = ob.OracleFDR(maxdrift=0)
    for c in range(3):
        blobs, _, _ = ob.demodulate(x[c], of.transform(x[c]))
        assert any(bytes(b) == bytes(meta["msg"]) for b in blobs)
