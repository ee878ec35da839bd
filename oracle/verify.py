"""End-to-end parity of the CUDA path against the reference on the SAME samples, at scale
(test infrastructure: used by tests/ and by bench.py's verification leg, after its timed
regions -- never on the product path).

For every window the reference chain is run on the host:

  * `oracle/_ref` (the reference's own FDR_impl.cc / sync_and_demodulate_impl.cc, unmodified,
    see oracle/Makefile) when it was built: FDR -> sync_and_demodulate with its own FFT stub,
    every sync_and_demodulate() call and every published message recorded by interposition;
  * else the plain-C restatement (oracle/uwspr_oracle.c), which is pinned bit for bit to it.

and compared with what the GPU returned for that window:

  cand_set_mismatch     windows whose candidate list differs in length or in any of
                        (freq, shift, drift model and its parameters); sync/snr are compared to
                        1e-4 relative (the FFT is the one unpinned piece: FFTW3f in the reference,
                        a double-precision DFT in the stub, a radix-8 Stockham FFT on the GPU)
  refined_mismatch      candidates of candidate-set-equal windows whose refined (f1, shift1, drift1)
                        or gate decision differ from the reference's call trace
                        (lib/sync_and_demodulate_impl.cc:404-456)
  soft_symbol_mismatch  mode-2 evaluations (:460-475) whose 162 soft symbols or sync differ
  message_mismatch      windows whose list of published 7-byte messages differs (:484-490)

Windows with a candidate-set difference are then re-examined: the oracle's post-FFT chain
(normalizer, peak pick, coarse search) is run on the GPU's own power spectrogram; if that
reproduces the GPU's list bit for bit the difference is attributed to FFT rounding alone
(`explained_by_fft_rounding`), and the sync values of the two competing choices are reported.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FL = 45000
NJIG = 17


def _cand_key(c):
    """(freq, shift, model, model parameters) of one candidate as a hashable tuple"""
    if int(c["m_type"]) == 0:
        return (float(c["freq"]), int(c["shift"]), 0, float(c["lin_drift"]))
    return (float(c["freq"]), int(c["shift"]), 1, float(c["V1"]), float(c["V2"]), int(c["p1"]), int(c["p2"]))


def cand_sets_equal(a, b, rtol=1e-4):
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if _cand_key(x) != _cand_key(y):
            return False
        if abs(float(x["sync"]) - float(y["sync"])) > rtol * abs(float(y["sync"])) + 1e-30:
            return False
        if abs(float(x["snr"]) - float(y["snr"])) > rtol * max(1.0, abs(float(y["snr"]))):
            return False
    return True


def _split_calls(cands, calls):
    """groups the recorded sync_and_demodulate() calls by candidate (demodulate() :404-482):
    A (mode 0), B (mode 1), C twice (mode 1, linear model only), and when gated D (mode 0, mode 1)
    followed by 1..17 mode-2 calls.  Returns a list of (chain_calls, mode2_calls)."""
    out, k, n = [], 0, len(calls)
    for c in cands:
        head = 2 + (2 if int(c["m_type"]) == 0 else 0)
        chain = list(calls[k:k + head])
        k += head
        m2 = []
        if k + 1 < n + 1 and k < n and calls[k].mode == 0 and k + 1 < n and calls[k + 1].mode == 1 and _is_fine_grid(calls[k]):
            chain += [calls[k], calls[k + 1]]
            k += 2
            while k < n and calls[k].mode == 2:
                m2.append(calls[k])
                k += 1
        out.append((chain, m2))
    assert k == n, (k, n)
    return out


def _is_fine_grid(call):
    return call.lagstep == 16      # D's lag search (:445); A uses 64 (:411)


def _worker(job):
    (xpath, stride, lo, hi, rpath, params, use_ref, full_upto) = job
    sys.path.insert(0, ROOT)
    from oracle import port_binding as ob
    stream = np.load(xpath, mmap_mode="r")
    r = np.load(rpath, mmap_mode="r", allow_pickle=False)
    base, cands, refined, jig, soft = r["base"], r["cands"].view(ob.CAND_DTYPE).reshape(-1), r["refined"], r["jig"], r["soft"]
    msgs_flat, msg_base = r["msgs"], r["msg_base"]
    if use_ref:
        from oracle import ref_binding as rb
        fdr = rb.RefFDR(**params)
        sd = rb.RefSD(params["fs"], params["fl"], params["spb"], params["maxdrift"], params["maxfreqs"], params["cf"], logdir="/tmp")
    of = ob.OracleFDR(**params)
    res = dict(windows=0, cands=0, cand_set_mismatch=[], refined_mismatch=0, soft_symbol_mismatch=0, evaluations=0,
               message_mismatch=[], max_rel_sync_diff=0.0, full_jiggle_windows=0, details=[])
    for w in range(lo, hi):
        x = np.ascontiguousarray(stream[w * stride:w * stride + FL])
        g0, g1 = int(base[w]), int(base[w + 1])
        gc = cands[g0:g1]
        gm = [bytes(m) for m in msgs_flat[int(msg_base[w]):int(msg_base[w + 1])]]
        if use_ref:
            oc, blobs, calls, _ = rb.pipeline(fdr, sd, x)
        else:
            oc = of.transform(x)
            blobs, calls, _ = ob.demodulate(x, oc, cf=params["cf"])
        res["windows"] += 1
        res["cands"] += len(oc)
        om = [bytes(b) for b in blobs]
        if om != gm:
            res["message_mismatch"].append(w)
        if not cand_sets_equal(gc, oc):
            res["cand_set_mismatch"].append(w)
            continue
        for a, b in zip(gc, oc):
            if float(b["sync"]) != 0.0:
                res["max_rel_sync_diff"] = max(res["max_rel_sync_diff"], abs(float(a["sync"]) - float(b["sync"])) / abs(float(b["sync"])))
        per = _split_calls(oc, calls)
        for j, (chain, m2) in enumerate(per):
            g = g0 + j
            gated = len(chain) > 2 + (2 if int(oc[j]["m_type"]) == 0 else 0)
            bad = int(refined["worth_a_try"][g]) != int(gated)
            if gated and not bad:
                c0 = m2[0]
                bad = (np.float32(c0.f1_in).tobytes() != refined["f1"][g].tobytes()
                       or np.float32(c0.drift_in).tobytes() != refined["drift1"][g].tobytes()
                       or int(c0.shift_in) != int(refined["shift1"][g])
                       or np.float32(chain[-1].sync_out).tobytes() != refined["sync1"][g].tobytes())
            elif not bad:
                last = chain[-1]
                bad = (np.float32(last.f1_out).tobytes() != refined["f1"][g].tobytes() or int(last.shift_out) != int(refined["shift1"][g]))
            res["refined_mismatch"] += int(bad)
            for t, call in enumerate(m2):
                res["evaluations"] += 1
                if (int(call.shift_in) != int(jig["shift"][g, t]) or np.float32(call.sync_out).tobytes() != jig["sync"][g, t].tobytes()
                        or bytes(call.symbols) != soft[g, t].tobytes()):
                    res["soft_symbol_mismatch"] += 1
        if w < full_upto:
            # every one of the 17 jiggles (the reference stops at the first decode): the C restatement
            o_ref, o_jigs = ob.demodulate_full(x, oc, cf=params["cf"])
            res["full_jiggle_windows"] += 1
            for j in range(len(oc)):
                g = g0 + j
                for t, call in enumerate(o_jigs[j]):
                    res["evaluations"] += 1
                    if np.float32(call.sync_out).tobytes() != jig["sync"][g, t].tobytes() or bytes(call.symbols) != soft[g, t].tobytes():
                        res["soft_symbol_mismatch"] += 1
    return res


def oracle_on_gpu_ps(of, ctx, x, win=0):
    """the oracle's normalizer + peak pick + coarse search fed with the GPU's power spectrogram of
    window `win` of the last coarse call (kept bins spliced into the oracle's own spectrogram)"""
    ps = of.spectrogram(x)
    gps, gpsavg = ctx.debug_spectrogram(win)
    lo, nb = ctx.info.bin_lo, ctx.info.n_bins
    ps_o = ps[:, lo:lo + nb].copy()
    ps[:, lo:lo + nb] = gps
    c0, psavg, _ = of.normalize_peaks(ps)
    return of.coarse(ps, c0), ps_o, gps, psavg[lo:lo + nb], gpsavg


def cands_equal_exact(a, b):
    """every field identical; snr alone is compared to 4 ulp: it is 10*log10f(x), and libm's
    log10f is not correctly rounded (the value depends on the glibc version of the host the
    reference runs on), while the CUDA path rounds a double-precision log10"""
    from oracle import testdata as td
    if len(a) != len(b):
        return False
    ca, cb = td.canon_cands(a), td.canon_cands(b)
    if not np.allclose(ca["snr"], cb["snr"], rtol=5e-7, atol=1e-6):
        return False
    ca["snr"] = 0
    cb["snr"] = 0
    return ca.tobytes() == cb.tobytes()


def verify(stream, stride, nwin, params, npk, cands, refined, jig, soft, gpu_messages, cores=None, full_jiggle_windows=0,
           make_debug_context=None, max_detail=12):
    """stream: complex64 array holding window w at [w*stride, w*stride + 45000); npk..soft: what the CUDA path
    returned for these windows (17 jiggles per candidate); gpu_messages: per window, the list of 7-byte
    messages the product's host decoder published.  make_debug_context(nwin) -> a uwspr_b200 Context with
    set_debug(True) used to attribute candidate-set differences.  Returns the summary dict."""
    from oracle import port_binding as ob
    from oracle import ref_binding as rb
    use_ref = rb.available()
    cores = cores or len(os.sched_getaffinity(0))
    stream = np.ascontiguousarray(stream).reshape(-1)
    assert cands.dtype.itemsize == 48, "candidate records must keep the reference layout (48 bytes)"
    t0 = time.perf_counter()
    tag = "%d_%d" % (os.getpid(), int(t0 * 1e3) % 100000)
    xpath, rpath = "/dev/shm/uwspr_verify_x_%s.npy" % tag, "/dev/shm/uwspr_verify_r_%s.npz" % tag
    base = np.concatenate([[0], np.cumsum(npk)]).astype(np.int64)
    msg_base = np.concatenate([[0], np.cumsum([len(m) for m in gpu_messages])]).astype(np.int64)
    flat = [np.frombuffer(bytes(m), np.uint8) for ms in gpu_messages for m in ms]
    msgs = np.stack(flat) if flat else np.zeros((0, 7), np.uint8)
    try:
        np.save(xpath, stream[:(nwin - 1) * stride + FL])
        np.savez(rpath, base=base, cands=np.ascontiguousarray(cands).view(np.uint8), refined=np.ascontiguousarray(refined),
                 jig=np.ascontiguousarray(jig), soft=np.ascontiguousarray(soft), msgs=msgs, msg_base=msg_base)
        step = max(1, min(64, nwin // (4 * cores) or 1))
        jobs = [(xpath, stride, lo, min(nwin, lo + step), rpath, params, use_ref, full_jiggle_windows) for lo in range(0, nwin, step)]
        with mp.get_context("spawn").Pool(min(cores, len(jobs))) as pool:
            parts = pool.map(_worker, jobs, chunksize=1)
    finally:
        for p in (xpath, rpath):
            if os.path.exists(p):
                os.unlink(p)
    out = dict(windows=sum(p["windows"] for p in parts), candidates=sum(p["cands"] for p in parts),
               checker="oracle/_ref (unmodified reference sources)" if use_ref else "oracle port (C restatement)",
               cand_set_mismatch=sum(len(p["cand_set_mismatch"]) for p in parts),
               refined_mismatch=sum(p["refined_mismatch"] for p in parts),
               soft_symbol_mismatch=sum(p["soft_symbol_mismatch"] for p in parts),
               mode2_evaluations_compared=sum(p["evaluations"] for p in parts),
               full_jiggle_windows=sum(p["full_jiggle_windows"] for p in parts),
               message_mismatch=sum(len(p["message_mismatch"]) for p in parts),
               max_rel_sync_diff=max([p["max_rel_sync_diff"] for p in parts] + [0.0]), cores=cores)
    flips = sorted(w for p in parts for w in p["cand_set_mismatch"])
    out["message_mismatch_windows"] = sorted(w for p in parts for w in p["message_mismatch"])[:max_detail]
    out["cand_set_mismatch_windows"] = flips[:max_detail]
    # attribution of the candidate-set differences to FFT rounding
    if flips and make_debug_context is not None:
        of = ob.OracleFDR(**params)
        dctx = make_debug_context(len(flips))
        xs = np.stack([stream[w * stride:w * stride + FL] for w in flips])
        g_npk, g_cands = dctx.coarse(xs)
        gb = np.concatenate([[0], np.cumsum(g_npk)])
        explained, details = 0, []
        for i, w in enumerate(flips):
            want, *_ = oracle_on_gpu_ps(of, dctx, xs[i], i)
            mine = g_cands[gb[i]:gb[i + 1]]
            same_as_run = np.array_equal(td_canon(mine), td_canon(cands[base[w]:base[w + 1]]))
            ok = cands_equal_exact(mine, want) and same_as_run
            explained += int(ok)
            if len(details) < max_detail:
                oc = of.transform(xs[i])
                details.append(dict(window=int(w), explained=bool(ok), gpu=[_cand_key(c) + (float(c["sync"]),) for c in mine],
                                    reference=[_cand_key(c) + (float(c["sync"]),) for c in oc]))
        out["explained_by_fft_rounding"] = explained
        out["cand_set_mismatch_detail"] = details
    out["seconds"] = time.perf_counter() - t0
    return out


def td_canon(c):
    from oracle import testdata as td
    return td.canon_cands(c).view(np.uint8)
