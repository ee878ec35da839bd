#!/usr/bin/env python
"""SASS opcode histogram per kernel of libuwspr_b200.so (cuobjdump -sass), with the counts that matter for the
arithmetic contract and the north_star's staging claims spelled out:

  FFMA2 (addend = the -0.0 parameter)   packed products rounded once (common.cuh): NOT contractions
  FFMA2 (other addend)                  genuine packed fma where the product is exact (+-1 x d in k_coarse)
  FFMA / FMUL / FADD                    scalar fp32; FFMA only inside IEEE division / square-root sequences, the
                                        explicit fmaf(+-1, d, ss) of k_coarse's fifth bin, and the (unpinned) FFT
  UBLKCP / SYNCS                        TMA bulk copy + mbarrier;  LDGSTS  cp.async

Usage: tools/sass_histogram.py [path/to/libuwspr_b200.so] > profiles/r2_sass_histogram.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "gr-uwspr_b200/libuwspr_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist, notes = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name).split("(")[0]
        hist[kern], notes[kern] = collections.Counter(), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);", line)
    if not m or kern is None:
        continue
    op, args = m.group(1), m.group(2)
    base = op.split(".")[0]
    hist[kern][base] += 1
    if base == "FFMA2":
        last = args.split(",")[-1].strip()
        notes[kern]["FFMA2, addend in a uniform register / constant bank (the -0.0 pair)" if last.startswith(("UR", "c[")) else
                    "FFMA2, addend in a vector register"] += 1
    if base in ("UBLKCP", "LDGSTS", "SYNCS", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "DFMA"):
        notes[kern][base] += 1
for k in sorted(hist):
    total = sum(hist[k].values())
    print("== %s  (%d SASS instructions)" % (k, total))
    print("   " + ", ".join("%s %d" % (o, n) for o, n in hist[k].most_common(24)))
    for o, n in sorted(notes[k].items()):
        print("   * %-72s %d" % (o, n))
