#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export (SASS rows): top instructions by stall samples,
cumulative share by position, so that hot regions can be matched to the code."""
import csv
import sys


def main(path, top=30, window=0):
    rows = list(csv.reader(open(path)))
    h = rows[1]
    si, ie, ws = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    data = []
    for n, r in enumerate(rows[2:]):
        try:
            data.append((n, int(r[ws]), int(r[ie]), r[si].strip(), r))
        except (ValueError, IndexError):
            pass
    ts, ti = sum(d[1] for d in data), sum(d[2] for d in data)
    print(f"{rows[0][1]}: {len(data)} SASS instructions, {ti} warp-instructions executed, {ts} samples")
    tot = {c: 0 for _, c in stall_cols}
    for d in data:
        for i, c in stall_cols:
            tot[c] += int(d[4][i] or 0)
    print("stall reasons:", ", ".join(f"{c[6:]} {100 * v / ts:.1f}%" for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
    # share by tenths of the instruction stream
    nb = 20
    print("position histogram (samples% / inst%):")
    for b in range(nb):
        seg = data[b * len(data) // nb:(b + 1) * len(data) // nb]
        print(f"  [{seg[0][0]:5d}..{seg[-1][0]:5d}] {100 * sum(d[1] for d in seg) / ts:5.1f}% / {100 * sum(d[2] for d in seg) / ti:5.1f}%")
    print("top instructions:")
    for d in sorted(data, key=lambda d: -d[1])[:top]:
        print(f"  #{d[0]:5d} {100 * d[1] / ts:5.1f}% smp {100 * d[2] / ti:5.1f}% inst  {d[3][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
