"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors.

Everything here needs a B200 (`-m gpu`).  Bars:
  * spectrogram power: the FFT is the one unpinned piece (FFTW3f in the reference, a
    double-precision DFT in the oracle): |ps - ps_oracle| <= 2e-6 * max(ps row) per row;
  * everything after the FFT: bit exact.  The oracle's normalizer / coarse search is run on
    the GPU's own ps (kept bins spliced into the oracle's full spectrogram) and every field
    of every candidate must be identical, the sync metric included;
  * fine sync / soft symbols: bit exact against the oracle on the same candidates
    (f1, shift1, drift1, sync1, every jiggle's sync and all 162 soft symbols);
  * end to end on the fixtures: same candidates (bin, shift, drift model) as the golden
    reference run, sync within 1e-4 relative, decoded messages identical.
"""
import numpy as np
import pytest

from oracle import port_binding as ob
from oracle import testdata as td
from tests.conftest import case_window

pytestmark = pytest.mark.gpu

import uwspr_b200 as ub  # noqa: E402

CASES = ["ve3emb_c2", "test_1500", "rec_150613", "mix_whales", "syn_m0_w0", "syn_m0_w1", "syn_m0_w2", "syn_m0_w3",
         "syn_m0_w4", "syn_m0_w5", "syn_m4_w0", "syn_m4_w3", "syn_m4_w6", "syn_m4_w7"]

_ctx_cache = {}


def ctx_for(maxdrift=0, halfbandwidth=10, threshold=10, max_windows=64, **kw):
    key = (maxdrift, halfbandwidth, threshold, max_windows, tuple(sorted(kw.items())))
    if key not in _ctx_cache:
        _ctx_cache[key] = ub.Context(maxdrift=maxdrift, halfbandwidth=halfbandwidth, threshold=threshold,
                                     max_windows=max_windows, **kw)
    return _ctx_cache[key]


from oracle.verify import cands_equal_exact, oracle_on_gpu_ps  # noqa: E402


@pytest.mark.parametrize("name", CASES)
def test_coarse_against_oracle_and_golden(name, golden, golden_windows):
    md = int(golden["case_maxdrift"][CASES.index(name)])
    x = case_window(name, golden_windows)
    ctx = ctx_for(maxdrift=md)
    ctx.set_debug(True)
    npk, cands = ctx.coarse(x.reshape(1, -1))
    of = ob.OracleFDR(maxdrift=md)
    want, ps_o, ps_g, psavg_o, psavg_g = oracle_on_gpu_ps(of, ctx, x)
    # FFT-level agreement
    tol = 2e-6 * ps_o.max(axis=1, keepdims=True) + 1e-30
    assert np.all(np.abs(ps_g - ps_o) <= tol)
    # column sums are sequential fp32 sums of the GPU's own ps: bit exact
    assert np.array_equal(psavg_g, psavg_o)
    # post-FFT chain: bit exact
    assert npk[0] == len(want)
    assert cands_equal_exact(cands, want)
    # against the recorded reference run
    ref = golden[name + "/cands"].view(ob.CAND_DTYPE).reshape(-1)
    assert len(ref) == len(cands)
    for a, b in zip(cands, ref):
        assert a["freq"] == b["freq"] and a["shift"] == b["shift"] and a["m_type"] == b["m_type"]
        if a["m_type"] == 0:
            assert a["lin_drift"] == b["lin_drift"]
        else:
            assert (a["V1"], a["V2"], a["p1"], a["p2"]) == (b["V1"], b["V2"], b["p1"], b["p2"])
        assert abs(a["sync"] - b["sync"]) <= 1e-4 * abs(b["sync"])
        assert abs(a["snr"] - b["snr"]) <= 1e-4 * max(1.0, abs(b["snr"]))


def check_fine_against_oracle(ctx, x, cands, cf=1500):
    npk = np.array([len(cands)], np.int32)
    refined, jig, soft = ctx.fine(x.reshape(1, -1), npk, cands)
    o_ref, o_jigs = ob.demodulate_full(x, cands, cf=cf)
    for g in range(len(cands)):
        assert refined["f1"][g].tobytes() == o_ref[g, 0].tobytes()
        assert refined["shift1"][g] == int(o_ref[g, 1])
        assert refined["drift1"][g].tobytes() == o_ref[g, 2].tobytes()
        assert refined["sync1"][g].tobytes() == o_ref[g, 3].tobytes()
        assert refined["worth_a_try"][g] == int(o_ref[g, 4])
        if not refined["worth_a_try"][g]:
            assert not soft[g].any() and not jig["gate"][g].any()
            continue
        for t, call in enumerate(o_jigs[g]):
            assert jig["shift"][g, t] == call.shift_in
            assert jig["sync"][g, t].tobytes() == np.float32(call.sync_out).tobytes()
            assert np.array_equal(soft[g, t], np.frombuffer(bytes(call.symbols), np.uint8))
            y = soft[g, t].astype(np.float32) - 128.0
            rms = np.float32(np.sqrt(np.float64(np.float32((y * y).sum())) / 162.0))
            assert jig["rms"][g, t] == rms
            assert jig["gate"][g, t] == int(call.sync_out > np.float32(0.12) and rms > np.float32(40.625))
    return refined, jig, soft


@pytest.mark.parametrize("name", CASES)
def test_fine_bit_exact_and_messages(name, golden, golden_windows):
    md = int(golden["case_maxdrift"][CASES.index(name)])
    x = case_window(name, golden_windows)
    ctx = ctx_for(maxdrift=md)
    ref_cands = golden[name + "/cands"].view(ob.CAND_DTYPE).reshape(-1)
    refined, jig, soft = check_fine_against_oracle(ctx, x, ref_cands)
    msgs = ub.decode_candidates(refined, jig, soft)
    got = np.array([m for _, m, _ in msgs], np.uint8).reshape(-1, 7)
    assert got.tobytes() == golden[name + "/blobs"].tobytes()
    # the trace of the reference run: the last mode-2 call it made before decoding is one of ours
    gcalls = golden[name + "/calls"]
    mode2 = [c for c in gcalls if int(np.frombuffer(c[:4].tobytes(), np.int32)[0]) == 2]
    for c in mode2:
        sym = c[52:52 + 162]
        assert any(np.array_equal(sym, soft[g, t]) for g in range(len(soft)) for t in range(soft.shape[1]))


def test_end_to_end_block_mirrors(golden, golden_windows):
    """FDR -> sync_and_demodulate as a flowgraph wires them, one window at a time"""
    fdr = ub.FDR(375, 45000, 256, 0, 200, 10, 1500, 10)
    sd = ub.sync_and_demodulate(375, 45000, 256, 0, 200, 1500)
    for name in ["ve3emb_c2", "test_1500", "rec_150613", "mix_whales"]:
        x = golden_windows[name]
        cands = fdr.transform(x)
        msgs = sd.demodulate(x, cands)
        got = np.array(msgs, np.uint8).reshape(-1, 7)
        assert got.tobytes() == golden[name + "/blobs"].tobytes()


def test_nonlinear_candidates_and_intended_t(golden_windows):
    """both drift models through the fine stage, including hand-made nonlinear candidates"""
    x = golden_windows["ve3emb_c2"]
    ctx = ctx_for()
    cands = np.zeros(3, ub.CAND_DTYPE)
    for g, k in enumerate([0, 57, 124]):
        cands[g]["freq"], cands[g]["sync"], cands[g]["shift"], cands[g]["m_type"] = -0.7324219, 0.5, 256, 1
        cands[g]["V1"], cands[g]["V2"], cands[g]["p1"], cands[g]["p2"] = ob.slm_trajectory(k)
    check_fine_against_oracle(ctx, x, cands)
    ctx2 = ctx_for(nonlinear_intended_t=1)
    ob.lib().orc_set_nonlinear_intended_t(1)
    try:
        check_fine_against_oracle(ctx2, x, cands)
    finally:
        ob.lib().orc_set_nonlinear_intended_t(0)


def test_edges_shift_off_both_ends():
    """lags that run off the start and the end of the buffer (n <= 0 and n >= 45000 are skipped)"""
    x, _ = td.synth_window(21, 0, snr_db=-12.0, start=375)
    ctx = ctx_for(maxdrift=4)
    cands = np.zeros(4, ub.CAND_DTYPE)
    for g, (sh, dr) in enumerate([(-300, 0.0), (0, 1.0), (3200, -2.0), (3700, 3.0)]):
        cands[g]["freq"], cands[g]["sync"], cands[g]["shift"], cands[g]["m_type"], cands[g]["lin_drift"] = 1.4648438, 0.3, sh, 0, dr
    check_fine_against_oracle(ctx, x, cands)


def test_batch_matches_per_window_oracle():
    """32 seeded synthetic windows, maxdrift 4, SNR -30..0 dB, one submission"""
    nwin = 32
    xs, metas = td.synth_batch(nwin, stream=2)
    ctx = ctx_for(maxdrift=4)
    ctx.set_debug(True)
    npk, cands, refined, jig, soft = ctx.coarse_fine(xs)
    of = ob.OracleFDR(maxdrift=4)
    base = np.concatenate([[0], np.cumsum(npk)])
    decoded = 0
    for w in range(nwin):
        want, *_ = oracle_on_gpu_ps(of, ctx, xs[w], w)
        got = cands[base[w]:base[w + 1]]
        assert cands_equal_exact(got, want)
        o_ref, o_jigs = ob.demodulate_full(xs[w], want)
        for j in range(len(want)):
            g = base[w] + j
            assert refined["f1"][g].tobytes() == o_ref[j, 0].tobytes() and refined["shift1"][g] == int(o_ref[j, 1])
            assert refined["sync1"][g].tobytes() == o_ref[j, 3].tobytes()
            for t, call in enumerate(o_jigs[j]):
                assert np.array_equal(soft[g, t], np.frombuffer(bytes(call.symbols), np.uint8))
        msgs = ub.decode_candidates(refined[base[w]:base[w + 1]], jig[base[w]:base[w + 1]], soft[base[w]:base[w + 1]])
        oblobs, _, _ = ob.demodulate(xs[w], want)
        assert np.array([m for _, m, _ in msgs], np.uint8).reshape(-1, 7).tobytes() == oblobs.tobytes()
        decoded += any(np.array_equal(m, metas[w]["msg"]) for _, m, _ in msgs)
    assert decoded >= nwin // 2  # most of U(-30, 0) dB decodes


def test_chunking_overlap_and_device_pointer():
    """results do not depend on chunk size, on window overlap (stride < fl) or on where samples live"""
    import torch
    nwin, stride = 9, 22500
    stream = np.concatenate([td.synth_window(4, w, snr_db=-14.0)[0][:stride] for w in range(nwin + 1)])
    windows = np.stack([stream[w * stride:w * stride + 45000] for w in range(nwin)])
    big = ctx_for(max_windows=64)
    a = big.coarse_fine(windows)
    b = big.coarse_fine(stream, nwin=nwin, stride=stride)
    small = ub.Context(max_windows=16, max_candidates=64)
    small.info.max_windows  # noqa: B018
    t = torch.from_numpy(stream.view(np.float32)).cuda()
    c = big.coarse_fine((t.data_ptr(), stream.size), nwin=nwin, stride=stride)
    for r in (b, c):
        for u, v in zip(a, r):
            assert u.tobytes() == v.tobytes()
    # chunk size 2048 is fixed inside the library; force several chunks with a tiny context
    tiny = ub.Context(max_windows=4096, max_candidates=8192)
    many = np.concatenate([windows] * 300)[:2050]
    npk, cands, refined, jig, soft = tiny.coarse_fine(many)
    assert npk.sum() == len(cands)
    k = int(a[0].sum())
    assert cands[:k].tobytes() == a[1].tobytes() and soft[:k].tobytes() == a[4].tobytes()
    assert np.array_equal(npk[:nwin], npk[nwin:2 * nwin]) and np.array_equal(npk[2043:2050], npk[0:7])
    tiny.close()
    small.close()


def test_empty_noise_and_zero_windows():
    ctx = ctx_for()
    z = np.zeros((2, 45000), np.complex64)
    npk, cands = ctx.coarse(z)
    of = ob.OracleFDR()
    assert list(npk) == [len(of.transform(z[0]))] * 2
    n, _ = td.synth_window(8, 0, snr_db=-80.0)
    npk, cands, refined, jig, soft = ctx.coarse_fine(n.reshape(1, -1))
    assert npk[0] == len(of.transform(n))
    assert ctx.coarse_fine(np.zeros((0, 45000), np.complex64), nwin=0, fetch=False) == 0


def test_wide_band_many_candidates():
    """halfbandwidth 100: 274 pass-band bins, several signals, two-digit candidate counts"""
    x = sum(td.synth_window(9, w, snr_db=-10.0 - w, f0=f0)[0] for w, f0 in enumerate([-80.0, -33.0, 4.0, 41.0, 77.0]))
    ctx = ub.Context(halfbandwidth=100, maxdrift=1, max_windows=2)
    ctx.set_debug(True)
    npk, cands = ctx.coarse(np.stack([x, x[::-1].copy()]))
    of = ob.OracleFDR(halfbandwidth=100, maxdrift=1)
    want, *_ = oracle_on_gpu_ps(of, ctx, x, 0)
    assert npk[0] == len(want) >= 5
    assert cands_equal_exact(cands[:npk[0]], want)
    check_fine_against_oracle(ctx, x, want[:6])
    ctx.close()


def test_errors_are_statuses_not_exits():
    with pytest.raises(ub.UwsprError) as e:
        ub.Context(halfbandwidth=188)          # the reference exits (FDR_impl.cc:85-90)
    assert e.value.status == 1
    with pytest.raises(ub.UwsprError):
        ub.Context(halfbandwidth=187)          # the reference reads out of bounds
    with pytest.raises(ub.UwsprError):
        ub.Context(spb=128)
    ctx = ub.Context(max_windows=4, max_candidates=1, halfbandwidth=40)
    a = td.synth_window(3, 0, snr_db=-12.0, f0=-20.0)[0] + td.synth_window(3, 1, snr_db=-15.0, f0=17.0)[0]
    with pytest.raises(ub.UwsprError) as e:
        ctx.coarse(a.reshape(1, -1))
    assert e.value.status == 3                  # capacity
    with pytest.raises(ub.UwsprError):
        ctx.coarse(np.zeros((5, 45000), np.complex64))  # nwin > max_windows
    ctx.close()


def test_batched_receiver_on_sliding_stream():
    """uwspr_b200_receiver: window k = stream[k*shift*fs, +fl) (sliding_window_stream_to_pdu_impl.cc:113-135),
    batches of 5 windows per submission; per window the same messages, in the same order, as the
    reference chain FDR -> sync_and_demodulate produces for that window"""
    from oracle import testdata as tdd
    stride, nwin = 9 * 375, 13
    n = 45000 + (nwin - 1) * stride
    rng = np.random.default_rng(77)
    stream = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * np.sqrt(0.15 / 10 ** (-1.2) / 2.0)
    truth = []
    for start, f0, seed in ((9000, -4.0, 1), (37000, 3.1, 2)):
        msg = tdd.message_bytes(np.random.default_rng(seed))
        sig = tdd.modulate(ob.channel_symbols(msg), f0=f0, drift=0.0, start=0, fl=162 * 256)
        stream[start:start + 162 * 256] += sig
        truth.append(bytes(msg))
    stream = stream.astype(np.complex64)
    rx = ub.Receiver(maxdrift=0, shift=9, batch_windows=5)
    got = []
    for lo in range(0, n, 10000):   # arbitrary push sizes
        got += rx.push(stream[lo:lo + 10000])
    got += rx.push(np.zeros(0, np.complex64), flush=True)
    assert rx.windows_done() == nwin
    of = ob.OracleFDR(maxdrift=0)
    want = []
    for k in range(nwin):
        x = stream[k * stride:k * stride + 45000]
        oc = of.transform(x)
        blobs, _, _ = ob.demodulate(x, oc)
        want += [(k, bytes(b)) for b in blobs]
    assert [(w, bytes(m)) for w, m, _ in got] == want
    assert {m for _, m in want} == set(truth)  # both transmissions are decoded (several times each)
    rx.close()


def test_parity_at_scale_256_windows():
    """256 seeded windows (maxdrift 4, SNR -30..0 dB) in one submission: every candidate record is
    identical to the oracle's post-FFT chain on the GPU's own spectrogram, every refinement result,
    every jiggle's sync and every soft symbol (256 x ~39 evaluations x 162 symbols) is identical to the
    oracle's, and the decoded messages equal the oracle's; >= 90 % of the payloads are recovered"""
    nwin = 256
    xs, metas = td.synth_batch(nwin, stream=31)
    ctx = ub.Context(maxdrift=4, max_windows=nwin)
    ctx.set_debug(True)
    npk, cands, refined, jig, soft = ctx.coarse_fine(xs)
    of = ob.OracleFDR(maxdrift=4)
    base = np.concatenate([[0], np.cumsum(npk)])
    recovered = 0
    for w in range(nwin):
        want, *_ = oracle_on_gpu_ps(of, ctx, xs[w], w)
        got = cands[base[w]:base[w + 1]]
        assert cands_equal_exact(got, want), w
        o_ref, o_jigs = ob.demodulate_full(xs[w], got)
        for j in range(len(got)):
            g = base[w] + j
            assert refined["f1"][g].tobytes() == o_ref[j, 0].tobytes() and refined["shift1"][g] == int(o_ref[j, 1]), w
            assert refined["drift1"][g].tobytes() == o_ref[j, 2].tobytes(), w
            assert refined["sync1"][g].tobytes() == o_ref[j, 3].tobytes(), w
            for t, call in enumerate(o_jigs[j]):
                assert jig["sync"][g, t].tobytes() == np.float32(call.sync_out).tobytes(), (w, t)
                assert soft[g, t].tobytes() == bytes(call.symbols), (w, t)
        msgs = ub.decode_candidates(refined[base[w]:base[w + 1]], jig[base[w]:base[w + 1]], soft[base[w]:base[w + 1]])
        recovered += any(bytes(m) == bytes(metas[w]["msg"]) for _, m, _ in msgs)
    assert recovered >= 0.9 * nwin
    ctx.close()


@pytest.mark.parametrize("maxfreqs,threshold,maxdrift", [(3, 10, 0), (200, 1, 2), (200, 1000000, 1), (2, 3, 4)])
def test_candidate_cap_and_threshold_variants(maxfreqs, threshold, maxdrift):
    """maxfreqs caps the peak list in bin order before the snr sort (FDR_impl.cc:296-305); the
    nonlinear/linear threshold changes which hypothesis the ordered update rule keeps (:360,:392)"""
    x = sum(td.synth_window(13, w, snr_db=-9.0 - 2 * w, f0=f0)[0] for w, f0 in enumerate([-31.0, -12.0, 6.5, 29.0]))
    kw = dict(halfbandwidth=40, maxfreqs=maxfreqs, threshold=threshold, maxdrift=maxdrift)
    ctx = ub.Context(max_windows=1, **kw)
    ctx.set_debug(True)
    npk, cands = ctx.coarse(x.reshape(1, -1))
    of = ob.OracleFDR(**kw)
    want, *_ = oracle_on_gpu_ps(of, ctx, x, 0)
    assert npk[0] == len(want) == min(maxfreqs, len(want))
    assert cands_equal_exact(cands, want)
    if threshold == 1:
        assert (cands["m_type"] == 1).any()      # a ratio of 1 lets trajectories displace the linear hits
    if threshold == 1000000:
        assert (cands["m_type"] == 0).all()
    check_fine_against_oracle(ctx, x, want[:3])
    ctx.close()


def test_staged_jiggles_equal_the_full_set(golden_windows):
    """fine(jig 0..0) followed by fine(jig 1..16) for the candidates that did not decode
    returns exactly the soft symbols of the eager 17-jiggle call; odd ranges exercise the generic lag loop"""
    ctx = ctx_for(maxdrift=4)
    xs = np.stack([golden_windows["mix_whales"], td.synth_window(0, 4, snr_db=-29.0)[0], td.synth_window(0, 7, snr_db=-23.0)[0]])
    npk, cands = ctx.coarse(xs)
    full = ctx.fine(xs, npk, cands)
    first = ctx.fine(xs, npk, cands, jig_first=0, jig_count=1)
    rest = ctx.fine(xs, npk, cands, jig_first=1, jig_count=16)
    mid = ctx.fine(xs, npk, cands, jig_first=5, jig_count=7)
    for k in range(3):  # refined, jig, soft
        if k == 0:
            assert full[0].tobytes() == first[0].tobytes() == rest[0].tobytes() == mid[0].tobytes()
            continue
        assert np.array_equal(full[k][:, 0:1], first[k])
        assert np.array_equal(full[k][:, 1:17], rest[k])
        assert np.array_equal(full[k][:, 5:12], mid[k])
    # and the candidate list left on the device by coarse() is the one fine(cands=None) uses
    again = ctx.fine(xs, jig_first=0, jig_count=17)
    t = len(cands)
    assert again[2][:t].tobytes() == full[2].tobytes()


def test_hydrophone_array_with_whale_noise():
    """BASELINE.json configs[4] at test size: 16 channels x 2 windows of a synthetic hydrophone array (one frame
    per window, per-channel delay and gain, independent noise, looped whale recording as interference), flattened
    (channel, window) -> one submission; every channel bit-exact against the oracle chain"""
    whales = np.load(td.golden_path("whales_375sps.npy"))
    nchan, nwin = 16, 2
    xs, metas = [], []
    for w in range(nwin):
        x, meta = td.synth_array(nchan, w, whales, snr_db=-21.0)
        xs.append(x)
        metas.append(meta)
    flat = np.stack(xs, axis=1).reshape(nchan * nwin, -1)   # index = channel * nwin + window
    ctx = ctx_for(maxdrift=0, max_windows=nchan * nwin)
    ctx.set_debug(True)
    npk, cands, refined, jig, soft = ctx.coarse_fine(flat)
    base = np.concatenate([[0], np.cumsum(npk)])
    of = ob.OracleFDR(maxdrift=0)
    heard = np.zeros((nchan, nwin), bool)
    for i in range(nchan * nwin):
        c, w = divmod(i, nwin)
        want, *_ = oracle_on_gpu_ps(of, ctx, flat[i], i)
        got = cands[base[i]:base[i + 1]]
        assert cands_equal_exact(got, want)
        o_ref, o_jigs = ob.demodulate_full(flat[i], want)
        for j in range(len(want)):
            g = base[i] + j
            assert refined["f1"][g].tobytes() == o_ref[j, 0].tobytes() and refined["shift1"][g] == int(o_ref[j, 1])
            for t, call in enumerate(o_jigs[j]):
                assert np.array_equal(soft[g, t], np.frombuffer(bytes(call.symbols), np.uint8))
        sl = slice(base[i], base[i + 1])
        msgs = ub.decode_candidates(refined[sl], jig[sl], soft[sl])
        oblobs, _, _ = ob.demodulate(flat[i], want)
        assert np.array([m for _, m, _ in msgs], np.uint8).reshape(-1, 7).tobytes() == oblobs.tobytes()
        heard[c, w] = any(np.array_equal(m, metas[w]["msg"]) for _, m, _ in msgs)
        # the refined shift follows the channel's delay to within a few of the 16-sample lag steps searched
        if heard[c, w]:
            j = [k for k, (_, m, _) in enumerate(msgs) if np.array_equal(m, metas[w]["msg"])][0]
            g = base[i] + msgs[j][0]
            assert abs(int(refined["shift1"][g]) - (metas[w]["start"] + int(metas[w]["delays"][c]))) <= 48
    assert heard.mean() > 0.7


def test_frontend_audio_to_messages():
    """12 kHz real audio -> GPU front-end (mix down 1500 Hz, FIR, keep every 32nd) -> 375-sps complex that stays on
    the device -> coarse + fine -> host decoder -> unpacker: the transmitted texts come back.  The front-end itself
    is checked against a float64 evaluation of the same formula (no GNU Radio here to pin its stock blocks)."""
    import torch
    from scipy.signal import fftconvolve
    texts = [("VE3EMB", "FN25", 30), ("K1ABC", "FN42", 37)]
    fs_in, decim, n_in = 12000.0, 32, 45000 * 32
    rng = np.random.default_rng(12)
    t = np.arange(n_in) / fs_in
    audio = np.zeros((2, n_in))
    for c, (call, grid, dbm) in enumerate(texts):
        sym = ub.channel_symbols(ub.pack_type1(call, grid, dbm)).astype(np.float64)
        # 4-FSK at audio: tone (sym - 1.5) * 375/256 Hz around 1500 + f0, 8192 audio samples per symbol
        f0 = (-3.3, 4.1)[c]
        start = int((1.0 + 0.37 * c) * fs_in)
        f = 1500.0 + f0 + (np.repeat(sym, 8192) - 1.5) * 375.0 / 256.0
        ph = 2 * np.pi * np.cumsum(f) / fs_in
        audio[c, start:start + 162 * 8192] = 0.2 * np.cos(ph)
        audio[c] += 0.05 * rng.standard_normal(n_in)
    taps = ub.lowpass_taps()
    # float64 oracle of the documented formula
    k = (len(taps) - 1) // 2
    want = np.empty((2, n_in // decim), np.complex128)
    for c in range(2):
        bb = audio[c] * np.exp(-2j * np.pi * 1500.0 * np.arange(n_in) / fs_in)
        full = fftconvolve(bb, taps.astype(np.float64))[k:][:n_in]
        want[c] = full[::decim]
    got = ub.frontend(audio.astype(np.float32), taps=taps)
    scale = np.sqrt((np.abs(want) ** 2).mean())
    assert got.shape == want.shape and np.abs(got - want).max() < 2e-5 * scale * 10
    pcm = np.clip(np.round(audio * 32768.0), -32768, 32767).astype(np.int16)
    got16 = ub.frontend(pcm, taps=taps)
    want16 = np.empty_like(want)
    for c in range(2):
        bb = (pcm[c] / 32768.0) * np.exp(-2j * np.pi * 1500.0 * np.arange(n_in) / fs_in)
        want16[c] = fftconvolve(bb, taps.astype(np.float64))[k:][:n_in][::decim]
    assert np.abs(got16 - want16).max() < 2e-4 * scale
    # device to device: audio and the 375-sps stream never leave the GPU
    a_dev = torch.from_numpy(audio.astype(np.float32)).cuda()
    x_dev = torch.empty((2, 45000, 2), dtype=torch.float32, device="cuda")
    import ctypes
    n_out = ub.frontend((ctypes.c_void_p(a_dev.data_ptr()), np.float32, tuple(a_dev.shape)), taps=taps,
                        out_device_ptr=x_dev.data_ptr())
    assert n_out == 45000
    torch.cuda.synchronize()
    assert np.array_equal(x_dev.cpu().numpy().view(np.complex64)[..., 0], got)
    ctx = ctx_for(maxdrift=0, max_windows=2)
    npk, cands, refined, jig, soft = ctx.coarse_fine((x_dev.data_ptr(), 2 * 45000), nwin=2, stride=45000)
    u = ub.WSPR_unpacker()
    base = np.concatenate([[0], np.cumsum(npk)])
    for c, (call, grid, dbm) in enumerate(texts):
        sl = slice(base[c], base[c + 1])
        heard = {u.unpack(m)[1] for _, m, _ in ub.decode_candidates(refined[sl], jig[sl], soft[sl])}
        assert "%s %s %2d" % (call, grid, dbm) in heard
    # and the same windows through the oracle chain decode to the same blobs
    of = ob.OracleFDR(maxdrift=0)
    for c in range(2):
        oc = of.transform(got[c])
        blobs, _, _ = ob.demodulate(got[c], oc)
        sl = slice(base[c], base[c + 1])
        mine = [bytes(m) for _, m, _ in ub.decode_candidates(refined[sl], jig[sl], soft[sl])]
        assert mine == [bytes(b) for b in blobs]


def test_frontend_equals_the_flowgraph_cascade():
    """the front-end with the composite complex taps (flowgraph_taps, delay 0) against the block-by-block float64
    restatement of the flowgraph's GNU Radio blocks (oracle/gr_frontend.py: band-pass at centre 0 -> translation by
    1500 Hz + low-pass -> rational resampler 1/32; examples/WaveFilePlusNoiseDecode.grc:834-958,1753-1810), on two
    channels of WSPR audio in noise; then the 375-sps stream decodes to the transmitted texts.
    Tolerance 5e-6 of the output rms (measured 5..7e-7): float32 products and sums over 6831 taps here, float64 there (GNU Radio's own
    float32 blocks would not agree with either more closely)."""
    from oracle import gr_frontend as gf
    texts = [("VE3EMB", "FN25", 30), ("W1AW", "FN31", 23)]
    fs_in, n_in = 12000.0, 45000 * 32
    rng = np.random.default_rng(21)
    audio = np.zeros((2, n_in))
    for c, (call, grid, dbm) in enumerate(texts):
        sym = ub.channel_symbols(ub.pack_type1(call, grid, dbm)).astype(np.float64)
        f0 = (2.7, -5.2)[c]
        start = int((0.8 + 0.5 * c) * fs_in)
        f = 1500.0 + f0 + (np.repeat(sym, 8192) - 1.5) * 375.0 / 256.0
        audio[c, start:start + 162 * 8192] = 0.1 * np.cos(2 * np.pi * np.cumsum(f) / fs_in)
        audio[c] += 0.3 * rng.standard_normal(n_in) + 0.4 * np.cos(2 * np.pi * 700.0 * np.arange(n_in) / fs_in)
    audio32 = audio.astype(np.float32)
    taps = ub.flowgraph_taps()
    assert taps.dtype == np.complex64 and len(taps) == 2891 + 2891 + 1051 - 2
    got = ub.frontend(audio32, taps=taps, delay=0)
    assert got.shape == (2, 45000)
    for c in range(2):
        want = gf.flowgraph_frontend(audio32[c].astype(np.float64))
        scale = np.sqrt((np.abs(want) ** 2).mean())
        err = np.abs(got[c] - want).max()
        print("front-end channel %d: max |gpu - cascade| = %.3g of the output rms" % (c, err / scale))
        assert err < 5e-6 * scale
        # the 700 Hz interferer (amplitude 0.4, four times the signal) is gone
        assert np.abs(want).max() < 0.2
    ctx = ctx_for(maxdrift=0, max_windows=2)
    npk, cands, refined, jig, soft = ctx.coarse_fine(got)
    u = ub.WSPR_unpacker()
    base = np.concatenate([[0], np.cumsum(npk)])
    for c, (call, grid, dbm) in enumerate(texts):
        sl = slice(base[c], base[c + 1])
        heard = {u.unpack(m)[1] for _, m, _ in ub.decode_candidates(refined[sl], jig[sl], soft[sl])}
        assert "%s %s %2d" % (call, grid, dbm) in heard


@pytest.mark.parametrize("piece,groups,stride", [(3, 2, 45000), (1, 2, 22500), (5, 1, 45000), (2, 2, 3375)])
def test_host_fed_tail_pieces(piece, groups, stride, monkeypatch):
    """host-fed calls cut their last chunk groups into small pieces on three streams and return results chunk by
    chunk; with tiny piece sizes every path of that schedule runs on a 21-window call, and the results must equal
    the single-submission ones byte for byte (independent and overlapping windows)"""
    nwin = 21
    n = 45000 + (nwin - 1) * stride
    rng = np.random.default_rng(99)
    stream = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.5).astype(np.complex64)
    for k, start in enumerate(range(2000, n - 162 * 256, 60000)):
        msg = td.message_bytes(np.random.default_rng(k))
        stream[start:start + 162 * 256] += td.modulate(ob.channel_symbols(msg), f0=-5.0 + 1.7 * (k % 6), drift=0.0, start=0,
                                                       fl=162 * 256).astype(np.complex64)
    big = ctx_for(max_windows=64)
    import torch
    t = torch.from_numpy(stream.view(np.float32)).cuda()
    want = big.coarse_fine((t.data_ptr(), stream.size), nwin=nwin, stride=stride)     # one chunk, device input
    monkeypatch.setenv("UWSPR_B200_TAIL_PIECE", str(piece))
    monkeypatch.setenv("UWSPR_B200_TAIL_GROUPS", str(groups))
    small = ub.Context(max_windows=21, max_candidates=21 * 14)   # host chunk groups of 10, 10 and 1 windows
    got = small.coarse_fine(stream, nwin=nwin, stride=stride)
    assert want[0].sum() >= 3
    for u, v in zip(want, got):
        assert u.tobytes() == v.tobytes()
    out = small.result_buffers(nwin)
    got2 = small.coarse_fine(stream, nwin=nwin, stride=stride, out=out)
    for u, v in zip(want, got2):
        assert u.tobytes() == np.asarray(v).tobytes()
    # the two block-level calls on their own, chunked the same way: coarse, then fine with the caller's list
    npk, cands = small.coarse(stream, nwin=nwin, stride=stride)
    assert npk.tobytes() == want[0].tobytes() and cands.tobytes() == want[1].tobytes()
    fine = small.fine(stream, npk, cands, nwin=nwin, stride=stride)
    for u, v in zip(want[2:], fine):
        assert u.tobytes() == v.tobytes()
    small.close()


def test_async_submit_equals_the_blocking_call():
    """uwspr_b200_coarse_fine_submit / _poll / _wait: same bytes as the blocking call, the blocking entry points refuse
    to run while a submission is outstanding"""
    xs, _ = td.synth_batch(24, stream=52)
    ctx = ub.Context(maxdrift=4, max_windows=24)
    want = ctx.coarse_fine(xs)
    out = ctx.result_buffers(24)
    assert ctx.poll() == -1
    ctx.submit(xs, out)
    with pytest.raises(ub.UwsprError) as e:
        ctx.coarse(xs)
    assert e.value.status == 5
    got = ctx.wait()
    assert ctx.poll() == -1
    for u, v in zip(want, got):
        assert u.tobytes() == np.asarray(v).tobytes()
    ctx.close()


def test_first_jiggle_taken_from_the_chain_equals_its_evaluation(monkeypatch):
    """stage E answers jiggle 0 (the refined point itself, sync_and_demodulate_impl.cc:461-464 with idt == 0) from the
    tone magnitudes a stage A-D evaluation of that point left behind; with UWSPR_B200_NO_FINE_REUSE the lag kernel
    evaluates all 17 lags again: same bytes either way, for all jiggles and for jiggle 0 alone"""
    xs, _ = td.synth_batch(48, stream=77)
    ctx = ub.Context(maxdrift=4, max_windows=48)
    want = ctx.coarse_fine(xs)
    npk, cands = want[0], want[1]
    first = ctx.fine(xs, npk, cands, jig_first=0, jig_count=1)
    ctx.close()
    monkeypatch.setenv("UWSPR_B200_NO_FINE_REUSE", "1")   # read once, when the context is created
    ctx = ub.Context(maxdrift=4, max_windows=48)
    got = ctx.coarse_fine(xs)
    first_again = ctx.fine(xs, npk, cands, jig_first=0, jig_count=1)
    ctx.close()
    assert want[3]["gate"].any()
    for u, v in zip(want, got):
        assert u.tobytes() == np.asarray(v).tobytes()
    for u, v in zip(first, first_again):
        assert u.tobytes() == np.asarray(v).tobytes()
    assert first[1][:, 0].tobytes() == want[3][:, 0].tobytes() and first[2][:, 0].tobytes() == want[4][:, 0].tobytes()


def test_fine_rejects_inconsistent_candidate_lists():
    """the caller's npk / total are checked on the host before anything is uploaded"""
    x, _ = td.synth_window(0, 1, snr_db=-10.0)
    ctx = ctx_for(maxdrift=0, max_windows=64)
    npk, cands = ctx.coarse(np.stack([x, x]))
    assert npk.sum() == len(cands) >= 2
    for bad in (np.array([-1, len(cands) + 1], np.int32), np.array([len(cands), 1], np.int32)):
        with pytest.raises(ub.UwsprError) as e:
            ctx.fine(np.stack([x, x]), bad, cands)
        assert e.value.status == 1
    # and a failed call does not leave a candidate list behind for fine(cands=None)
    with pytest.raises(ub.UwsprError) as e:
        ctx.fine(np.stack([x, x]))
    assert e.value.status == 5
    r = ctx.fine(np.stack([x, x]), npk, cands)       # the context is still usable
    assert r[0]["worth_a_try"].any()


def test_parity_config3_slice_against_the_reference_chain():
    """1 024 windows of the bench generator (BASELINE.json configs[2] statistics, maxdrift 4) through coarse_fine and,
    on the SAME samples, through the reference chain on the host (oracle/verify.py: the unmodified reference where
    oracle/_ref is built): no refined / soft-symbol / message difference; candidate-set differences, if any, are all
    reproduced by the oracle's post-FFT chain on the GPU's own spectrogram (FFT rounding)"""
    import torch
    from oracle import verify as vf
    from uwspr_b200 import synth
    nwin = 1024
    params = dict(fs=375, fl=45000, spb=256, maxdrift=4, maxfreqs=200, halfbandwidth=10, cf=1500, threshold=10)
    xs_t, truth = synth.gen_frames(0, nwin, 0, torch.device("cuda", 0))
    xs = xs_t.cpu().numpy()
    ctx = ub.Context(max_windows=nwin, **params)
    npk, cands, refined, jig, soft = ctx.coarse_fine(xs)
    dec = ub.decode_candidates(refined, jig, soft)
    win = np.repeat(np.arange(nwin), npk)
    msgs = [[] for _ in range(nwin)]
    for g, m, _ in dec:
        msgs[int(win[g])].append(bytes(m))

    def dbg(n):
        c = ub.Context(max_windows=n, **params)
        c.set_debug(True)
        return c
    r = vf.verify(xs.reshape(-1), 45000, nwin, params, npk, cands, refined, jig, soft, msgs, full_jiggle_windows=128, make_debug_context=dbg)
    print(r)
    assert r["windows"] == nwin
    assert r["refined_mismatch"] == 0 and r["soft_symbol_mismatch"] == 0
    assert r["cand_set_mismatch"] == r.get("explained_by_fft_rounding", 0)
    assert r["message_mismatch"] <= r["cand_set_mismatch"]
    assert sum(bytes(truth[w]["msg"]) in msgs[w] for w in range(nwin)) >= 0.9 * nwin
    ctx.close()
