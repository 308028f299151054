#!/usr/bin/env python
"""Markdown summary of an `ncu --page raw --csv` export: one row per kernel launch with duration,
DRAM bytes, achieved DRAM GB/s (fraction of the measured copy peak), issue-slot utilisation, achieved
occupancy, registers and the top two warp-stall reasons (pc sampling).

    python scripts/ncu_summary.py gpurun_out/x_hot_raw.csv[.xz] [peak_gbs] > profiles/x.md
"""
import csv
import json
import lzma
import os
import sys


def main():
    path = sys.argv[1]
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
    if peak is None:
        try:
            peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
        except (OSError, KeyError, ValueError):
            peak = 6545.9
    f = lzma.open(path, "rt") if path.endswith(".xz") else open(path)
    rows = list(csv.reader(f))
    h, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}

    def get(r, name, scale_unit=None):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return None
        v = float(r[i].replace(",", ""))
        u = units[i]
        if scale_unit == "bytes":
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        if scale_unit == "ms":
            v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}.get(u, 1)
        return v

    stall_cols = [(n, i) for n, i in col.items() if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")]
    print(f"| kernel | ms | DRAM read GB | DRAM write GB | GB/s | of {peak:.0f} | issue % | occupancy % | regs | top stalls |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    total = 0.0
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mc::", "")
        ms = get(r, "gpu__time_duration.sum", "ms")
        rd, wr = get(r, "dram__bytes_read.sum", "bytes") or 0.0, get(r, "dram__bytes_write.sum", "bytes") or 0.0
        issue = get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        occ = get(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
        regs = get(r, "launch__registers_per_thread")
        st = []
        for n, i in stall_cols:
            try:
                st.append((float(r[i].replace(",", "")), n.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
        st.sort(reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st[:2])
        gbs = (rd + wr) / (ms * 1e-3) / 1e9 if ms else 0.0
        total += ms or 0.0
        print(f"| {name} | {ms:.3f} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | {gbs:.0f} | {gbs / peak:.2f} | {issue:.0f} | {occ:.0f} | {regs:.0f} | {top} |")
    print(f"\nsum of the listed launches: {total:.2f} ms (per-launch times under ncu are cold-cache and serialised)")


if __name__ == "__main__":
    main()
