#!/usr/bin/env python
"""Per-kernel DRAM traffic of one profiled step as JSON (bench.py copies it into `roofline_stuffing`):

    python scripts/ncu_traffic_json.py gpurun_out/x_hot_raw.csv.xz csg64_1024@1 profiles/r02_hot_kernels_ncu.md > profiles/r02_stuffing_traffic.json
"""
import csv
import json
import lzma
import sys


def main():
    path, key, source = sys.argv[1:4]
    f = lzma.open(path, "rt") if path.endswith(".xz") else open(path)
    rows = list(csv.reader(f))
    h, u = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}

    def g(r, n):
        v = float(r[col[n]].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3,
                    "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(u[col[n]], 1)

    kernels = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mc::", "")
        if name.startswith(("k_sdf_field", "k_tile_decide", "k_subbox_decide")):
            continue
        e = kernels.setdefault(name, {"launches": 0, "ms_under_ncu": 0.0, "dram_bytes_read": 0, "dram_bytes_write": 0})
        e["launches"] += 1
        e["ms_under_ncu"] = round(e["ms_under_ncu"] + g(r, "gpu__time_duration.sum"), 4)
        e["dram_bytes_read"] += int(g(r, "dram__bytes_read.sum"))
        e["dram_bytes_write"] += int(g(r, "dram__bytes_write.sum"))
    total = sum(e["dram_bytes_read"] + e["dram_bytes_write"] for e in kernels.values())
    json.dump({key: {"dram_bytes_per_step": total, "per_kernel": kernels,
                     "source": source + ": ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of every stuffing / "
                               "boundary / side-set kernel of one step (the scans and list kernels move < 10 MB and are not captured)"}},
              sys.stdout, indent=1)


if __name__ == "__main__":
    main()
