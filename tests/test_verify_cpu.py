"""The at-scale verifier (oracle/verify.py) checked on the CPU: fed with results produced by the
oracle's C restatement it reports no difference against the reference chain (the unmodified
reference where oracle/_ref is built), and it counts every kind of corruption."""
import numpy as np

from oracle import port_binding as ob
from oracle import testdata as td
from oracle import verify as vf

REFINED_DTYPE = np.dtype([("f1", "<f4"), ("shift1", "<i4"), ("drift1", "<f4"), ("sync1", "<f4"), ("worth_a_try", "<i4"), ("reserved", "<i4")])
JIG_DTYPE = np.dtype([("sync", "<f4"), ("rms", "<f4"), ("shift", "<i4"), ("gate", "<i4")])
PARAMS = dict(fs=375, fl=45000, spb=256, maxdrift=4, maxfreqs=200, halfbandwidth=10, cf=1500, threshold=10)


def results_from_port(xs, params):
    """what the CUDA path returns (compact candidate-major arrays, 17 jiggles each), computed by the port"""
    of = ob.OracleFDR(**params)
    npk, cands, refined, jig, soft, msgs = [], [], [], [], [], []
    for x in xs:
        c = of.transform(x)
        o_ref, o_jigs = ob.demodulate_full(x, c, cf=params["cf"])
        npk.append(len(c))
        cands.append(c)
        r = np.zeros(len(c), REFINED_DTYPE)
        j = np.zeros((len(c), 17), JIG_DTYPE)
        s = np.zeros((len(c), 17, 162), np.uint8)
        for g in range(len(c)):
            r[g] = (o_ref[g, 0], int(o_ref[g, 1]), o_ref[g, 2], o_ref[g, 3], int(o_ref[g, 4]), 0)
            for t, call in enumerate(o_jigs[g]):
                j[g, t]["sync"], j[g, t]["shift"] = call.sync_out, call.shift_in
                s[g, t] = np.frombuffer(bytes(call.symbols), np.uint8)
        refined.append(r)
        jig.append(j)
        soft.append(s)
        blobs, _, _ = ob.demodulate(x, c, cf=params["cf"])
        msgs.append([bytes(b) for b in blobs])
    flat = np.zeros(sum(npk), ob.CAND_DTYPE)     # np.concatenate would repack the overlapping-field dtype
    k = 0
    for c in cands:
        flat[k:k + len(c)] = c
        k += len(c)
    return (np.array(npk, np.int32), flat, np.concatenate(refined), np.concatenate(jig), np.concatenate(soft), msgs)


def test_verifier_accepts_exact_results_and_counts_corruptions():
    nwin = 6
    xs, _ = td.synth_batch(nwin, stream=41)
    xs[5] = td.synth_window(41, 5, snr_db=-60.0)[0]      # a window without a decodable frame
    npk, cands, refined, jig, soft, msgs = results_from_port(xs, PARAMS)
    r = vf.verify(xs.reshape(-1), 45000, nwin, PARAMS, npk, cands, refined, jig, soft, msgs, cores=2, full_jiggle_windows=2)
    assert r["windows"] == nwin and r["candidates"] == len(cands)
    assert (r["cand_set_mismatch"], r["refined_mismatch"], r["soft_symbol_mismatch"], r["message_mismatch"]) == (0, 0, 0, 0)
    assert r["mode2_evaluations_compared"] >= len(cands) and r["full_jiggle_windows"] == 2
    # corrupt one of each
    base = np.concatenate([[0], np.cumsum(npk)])
    gated = np.flatnonzero(refined["worth_a_try"])
    g = int(gated[0])
    soft2 = soft.copy()
    soft2[g, 0, 17] ^= 1
    refined2 = refined.copy()
    refined2["shift1"][int(gated[1])] += 1
    cands2 = cands.copy()
    cands2["shift"][base[3]] += 128
    msgs2 = [list(m) for m in msgs]
    msgs2[0] = msgs2[0] + [b"\x00" * 7]
    r = vf.verify(xs.reshape(-1), 45000, nwin, PARAMS, npk, cands2, refined2, jig, soft2, msgs2, cores=2)
    assert r["cand_set_mismatch"] == 1 and r["cand_set_mismatch_windows"] == [3]
    assert r["message_mismatch"] == 1 and r["message_mismatch_windows"] == [0]
    w_g = int(np.searchsorted(base, g, side="right") - 1)
    w_g1 = int(np.searchsorted(base, int(gated[1]), side="right") - 1)
    assert r["soft_symbol_mismatch"] == (0 if w_g == 3 else 1)
    assert r["refined_mismatch"] == (0 if w_g1 == 3 else 1)


def test_verifier_on_an_overlapped_stream():
    """windows every 22 500 samples of one stream (BASELINE.json configs[3] geometry)"""
    stride, nwin = 22500, 5
    stream = np.concatenate([td.synth_window(42, w, snr_db=-16.0)[0][:stride] for w in range(nwin + 1)])
    xs = np.stack([stream[w * stride:w * stride + 45000] for w in range(nwin)])
    npk, cands, refined, jig, soft, msgs = results_from_port(xs, PARAMS)
    r = vf.verify(stream, stride, nwin, PARAMS, npk, cands, refined, jig, soft, msgs, cores=2)
    assert (r["windows"], r["cand_set_mismatch"], r["refined_mismatch"], r["soft_symbol_mismatch"], r["message_mismatch"]) == (nwin, 0, 0, 0, 0)
