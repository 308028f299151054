#!/usr/bin/env python
"""Markdown table of an `ncu --metrics gpu__time_duration.sum --csv` launch list (one profiled step):

    python scripts/launch_list_md.py gpurun_out/x_launches.csv [bench.json] > profiles/x_launches.md
"""
import csv
import json
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    h = rows[0]
    col = {n: i for i, n in enumerate(h)}
    launches = []
    for r in rows[1:]:
        if len(r) < len(h) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        u = r[col["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mc::", "")
        launches.append((name, r[col["Grid Size"]], r[col["Block Size"]], ms))
    total = sum(l[3] for l in launches)
    print("| # | kernel | grid x block | ms | share |")
    print("|---|---|---|---|---|")
    for i, (name, grid, block, ms) in enumerate(launches):
        print(f"| {i} | `{name}` | {grid} x {block} | {ms:.3f} | {100 * ms / total:.1f} % |")
    print(f"\nSum of the {len(launches)} launches: {total:.2f} ms.")
    if len(sys.argv) > 2:
        d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
        print(f"Un-profiled step (CUDA events, same code): {d['ms_per_step']:.2f} ms; stage_ms {json.dumps(d['stage_ms'])}.")


if __name__ == "__main__":
    main()
