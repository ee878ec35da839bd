#!/usr/bin/env python
"""Summarises an .ncu-rep (read with `ncu -i ... --page raw --csv`): per kernel launch the duration,
issue utilisation, pipe activity, stall mix, DRAM bytes.
Usage: tools/ncu_summary.py report.ncu-rep [--traffic WINDOWS out.json]
--traffic also writes the DRAM bytes (read + write) per launch and their sum over the fine path's kernels
(k_fine_*) of the capture, which bench.py reports as roofline.traffic when run at WINDOWS windows."""
import csv
import subprocess
import sys

KEYS = [
    ("duration_ms", "gpu__time_duration.sum"), ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("pipe_fma_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("pipe_alu_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("pipe_fp64_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("inst_executed", "smsp__inst_executed.sum"), ("regs", "launch__registers_per_thread"),
    ("dyn_smem_B", "launch__shared_mem_per_block_dynamic"), ("grid", "launch__grid_size"),
    ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("smem_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed"),
]


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main(path, traffic=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    per_launch = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        if "dram__bytes_read.sum" in d:
            name = d.get("Kernel Name", "?").replace("<unnamed>::", "").replace("void ", "").split("(")[0]
            per_launch.append(dict(kernel=name, duration_ms=float(d["gpu__time_duration.sum"].replace(",", "")),
                                   dram_bytes=to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
                                   + to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])))
        print("== %s  (id %s)" % (d.get("Kernel Name", "?")[:60], d.get("ID")))
        for name, k in KEYS:
            if k in d:
                print("   %-24s %s %s" % (name, d[k], u.get(k, "")))
        st = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                     if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v), reverse=True)
        print("   stalls/issue: " + ", ".join("%s %.2f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], v) for v, k in st[:7]))


    if traffic:
        import json
        fine = sum(x["dram_bytes"] for x in per_launch if x["kernel"].startswith("k_fine"))
        json.dump(dict(windows=int(traffic[0]), fine_path_dram_bytes=fine, k_fine_dram_bytes=fine, launches=per_launch,
                       source=path.split("/")[-1] + " (ncu --set full, one launch of every heavy kernel of one step)"),
                  open(traffic[1], "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3:5] if len(sys.argv) >= 5 and sys.argv[2] == "--traffic" else None)
