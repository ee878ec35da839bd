"""k_fine_finish replaces the double division by 162 of the normalisation sums (sync_and_demodulate_impl.cc:243-244)
with a multiply and two fmas (uw_div162, fine.cu).  tools/div162_exhaustive.c compares that sequence with the IEEE
quotient for every finite fp32 input; this test builds and runs it (every bit pattern when the CPU has an fma unit,
every 61st otherwise: the software fma of libm is slow)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fma_sequence_equals_the_ieee_quotient_for_every_fp32(tmp_path):
    exe = str(tmp_path / "div162")
    subprocess.run(["gcc", "-O2", "-march=native", "-fopenmp", os.path.join(ROOT, "tools", "div162_exhaustive.c"), "-o", exe, "-lm"],
                   check=True)
    with open("/proc/cpuinfo") as f:
        has_fma = " fma " in f.read()
    out = subprocess.run([exe] + ([] if has_fma else ["61"]), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches" in out.stdout
    if has_fma:
        assert int(out.stdout.split()[1]) == (1 << 32) - (1 << 24)   # every finite fp32 value
