#!/usr/bin/env python
"""Per-source-line profile of one kernel: joins the SASS rows of an `ncu --page source --csv` export (warp
stall samples, instructions executed) with the line table of the shipped cubin (`nvdisasm --print-line-info`;
the library is built with -lineinfo), by instruction index.

    python scripts/ncu_line_profile.py gpurun_out/x_src_k.csv[.xz] 'k_tet_emit<unsigned int, (bool)1>' \\
        mc_mesher _ZN2mc10k_tet_emitIjLb1E [top]

(args: csv, substring of the demangled kernel name in the csv, cubin stem, prefix of the mangled name)."""
import collections
import csv
import lzma
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(stem, mangled):
    so = os.path.join(ROOT, "moosecad_b200", "libmoosecad_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith(stem) and f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout
    table, cur, on = [], ("?", 0), False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            on = ln[6:].startswith(mangled)
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            table.append(cur)
    return table


def main():
    path, kname, stem, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
    f = lzma.open(path, "rt") if path.endswith(".xz") else open(path)
    rows = list(csv.reader(f))
    # sections: a "Kernel Name" row, a header row, then one row per SASS instruction
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kname in r[1])
    h = rows[start + 1]
    si, ie, ws = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    stall = {c: i for i, c in enumerate(h) if c.startswith("stall_")}
    data = []
    for r in rows[start + 2:]:
        if r and r[0] == "Kernel Name":
            break
        try:
            data.append((int(r[ws]), int(r[ie]), r))
        except (ValueError, IndexError):
            pass
    table = line_table(stem, mangled)
    print(f"{rows[start][1][:100]}: {len(data)} SASS rows in the profile, {len(table)} in the cubin")
    if len(data) != len(table):
        print("WARNING: instruction counts differ - is this the library that was profiled?")
    ts, ti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for (smp, inst, r), key in zip(data, table):
        a = agg[key]
        a[0] += smp
        a[1] += inst
        for c, i in stall.items():
            try:
                a[2][c[6:]] += int(r[i] or 0)
            except ValueError:
                pass
    src_cache = {}

    def text(fn, ln):
        if fn not in src_cache:
            p = os.path.join(ROOT, "moosecad_b200", "csrc", fn)
            src_cache[fn] = open(p).read().splitlines() if os.path.exists(p) else []
        L = src_cache[fn]
        return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""

    print(f"total: {ts} samples, {ti / 1e9:.2f} G warp-instructions")
    for key, (smp, inst, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        tops = ", ".join(f"{k} {100 * v / max(smp, 1):.0f}%" for k, v in st.most_common(2))
        print(f"{100 * smp / ts:5.1f}% smp {100 * inst / ti:5.1f}% inst  {key[0]}:{key[1]:<5d} [{tops}]  {text(*key)}")


if __name__ == "__main__":
    main()
