"""The C++ mirrors of the reference blocks (gr-uwspr_b200/host/blocks.h) driven by a small C++
program, tests/cpp/test_blocks.cc: the sliding window on the CPU, the FDR -> sync_and_demodulate
chain on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gr-uwspr_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "build", "test_blocks")


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as ge
    ge.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = [os.path.join(ROOT, "tests", "cpp", "test_blocks.cc"), os.path.join(PKG, "host", "blocks.cc")]
    if not os.path.exists(EXE) or any(os.path.getmtime(s) > os.path.getmtime(EXE) for s in src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "host"),
                               *src, "-L" + PKG, "-luwspr_b200", "-Wl,-rpath," + PKG, "-o", EXE])
    return EXE


def test_sliding_window_block(exe):
    """window k == stream[k*shift*fs, k*shift*fs + fl), at most one PDU per work() call
    (lib/sliding_window_stream_to_pdu_impl.cc:98-138)"""
    out = subprocess.run([exe, "sliding"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip() == "windows 6 bad 0 multi 0"


@pytest.mark.gpu
@pytest.mark.parametrize("name,blob", [("ve3emb_c2", "d42c73eb3a7780"), ("rec_150613", "d42c73eb0d1840"), ("mix_whales", "d42c73eb3a7780")])
def test_fdr_to_sync_and_demodulate_chain(exe, name, blob):
    out = subprocess.run([exe, "window", os.path.join(ROOT, "tests", "golden", "win_%s.npy" % name)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert [l for l in lines if l.startswith("message")] == ["message " + blob]
    assert lines[-1] == "frames 1"
