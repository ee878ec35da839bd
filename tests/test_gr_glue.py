"""The compiled GNU Radio glue (gr-uwspr_b200/gr_glue: gr::block subclasses with the reference's factories,
message ports and PDU schemas, calling only the C ABI), built against the GNU Radio / PMT stand-ins of
oracle/stubs and driven by tests/cpp/test_gr_glue.cc.  The same window PDUs go through the reference's own
blocks (oracle/_ref where built, the golden vectors everywhere) and the published tuples / blobs are
compared field by field."""
import glob
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import port_binding as ob
from oracle import ref_binding as rb
from oracle import testdata as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gr-uwspr_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "build", "test_gr_glue")


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as ge
    ge.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = [os.path.join(ROOT, "tests", "cpp", "test_gr_glue.cc")] + sorted(glob.glob(os.path.join(PKG, "gr_glue", "lib", "*.cc")))
    deps = src + glob.glob(os.path.join(PKG, "gr_glue", "lib", "*.h")) + glob.glob(os.path.join(PKG, "gr_glue", "include", "uwspr", "*.h"))
    if not os.path.exists(EXE) or any(os.path.getmtime(s) > os.path.getmtime(EXE) for s in deps):
        # -std=gnu++11: what a GNU Radio 3.7 tree is built with
        subprocess.check_call(["g++", "-O2", "-std=gnu++11", "-Wall", "-I" + os.path.join(ROOT, "oracle", "stubs"),
                               "-I" + os.path.join(PKG, "gr_glue", "include"), "-I" + os.path.join(ROOT, "include"), *src,
                               "-L" + PKG, "-luwspr_b200", "-Wl,-rpath," + PKG, "-o", EXE])
    return EXE


def parse_dump(path):
    """[(candidates, blobs)] per window PDU"""
    raw = open(path, "rb").read()
    out, k = [], 0
    while k < len(raw):
        npk, = struct.unpack_from("<i", raw, k)
        k += 4
        c = np.frombuffer(raw, np.uint8, npk * 48, k).view(ob.CAND_DTYPE).copy()
        k += npk * 48
        nb, = struct.unpack_from("<i", raw, k)
        k += 4
        blobs = [raw[k + 7 * i:k + 7 * i + 7] for i in range(nb)]
        k += 7 * nb
        out.append((c, blobs))
    return out


def same_candidates(got, want):
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a["freq"] == b["freq"] and a["shift"] == b["shift"] and a["m_type"] == b["m_type"]
        if a["m_type"] == 0:
            assert a["lin_drift"] == b["lin_drift"]
        else:
            assert (a["V1"], a["V2"], a["p1"], a["p2"]) == (b["V1"], b["V2"], b["p1"], b["p2"])
        assert abs(a["sync"] - b["sync"]) <= 1e-4 * abs(b["sync"])       # the FFT is the one unpinned piece
        assert abs(a["snr"] - b["snr"]) <= 1e-4 * max(1.0, abs(b["snr"]))


def test_sliding_window_block(exe):
    out = subprocess.run([exe, "sliding"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.splitlines()[0] == "windows 6 bad 0 multi 0"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ve3emb_c2", "test_1500", "rec_150613", "mix_whales"])
def test_fdr_and_sync_blocks_publish_the_reference_pdus(exe, name, golden, tmp_path):
    win = os.path.join(ROOT, "tests", "golden", "win_%s.npy" % name)
    dump = str(tmp_path / "out.bin")
    out = subprocess.run([exe, "chain", win, dump], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0, out.stdout + out.stderr
    (cands, blobs), = parse_dump(dump)
    same_candidates(cands, golden[name + "/cands"].view(ob.CAND_DTYPE).reshape(-1))
    assert b"".join(blobs) == golden[name + "/blobs"].tobytes()
    log = open(tmp_path / "messagelog.txt").read()          # the block's side effect in its working directory
    assert log.startswith("Start time: ") and log.count("Frame: ") == len(blobs) and "Data: " + blobs[0].hex() in log
    if rb.available():
        x = np.load(win)
        fdr, sd = rb.RefFDR(), rb.RefSD(logdir=str(tmp_path))
        rc, rblobs, _, _ = rb.pipeline(fdr, sd, x)
        same_candidates(cands, rc)
        assert b"".join(blobs) == rblobs.tobytes()


@pytest.mark.gpu
def test_three_blocks_on_a_stream(exe, tmp_path):
    """sliding window (shift 9 s) -> FDR -> sync_and_demodulate on a stream holding one frame: every window's
    candidate tuples and blobs equal the reference chain's for that window"""
    stride, nwin = 9 * 375, 4
    n = 45000 + (nwin - 1) * stride
    rng = np.random.default_rng(5)
    stream = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * np.sqrt(0.15 / 10 ** (-1.0) / 2.0)
    msg = td.message_bytes(np.random.default_rng(3))
    stream[11000:11000 + 162 * 256] += td.modulate(ob.channel_symbols(msg), f0=2.2, drift=0.0, start=0, fl=162 * 256)
    stream = stream.astype(np.complex64)
    spath, dump = str(tmp_path / "stream.npy"), str(tmp_path / "out.bin")
    np.save(spath, stream)
    out = subprocess.run([exe, "stream", spath, "9", dump], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0 and out.stdout.strip() == "stream rc 0 windows %d" % nwin, out.stdout + out.stderr
    got = parse_dump(dump)
    of = ob.OracleFDR(maxdrift=0)
    heard = 0
    for k, (cands, blobs) in enumerate(got):
        x = stream[k * stride:k * stride + 45000]
        oc = of.transform(x)
        same_candidates(cands, oc)
        oblobs, _, _ = ob.demodulate(x, oc)
        assert b"".join(blobs) == oblobs.tobytes()
        heard += bytes(msg) in blobs
    assert heard >= 1     # the coarse search looks for a frame start within the first 3 328 samples of a window
