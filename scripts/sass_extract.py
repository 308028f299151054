#!/usr/bin/env python
"""SASS evidence for profiles/: opcode counts of the hot kernels of the shipped library (`cuobjdump -sass`),
the TMA / mbarrier instructions of k_warp_codes with their neighbourhood, the FP64 probe (DADD / DMUL only).

    python scripts/sass_extract.py > profiles/r02_sass_extract.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "moosecad_b200", "libmoosecad_b200.so")
WATCH = ("UTMALDG", "UTMASTG", "SYNCS", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "DADD", "DMUL", "DFMA", "DSETP", "MUFU",
         "FFMA", "FADD", "FMUL", "BAR", "SHFL", "VOTE")
KERNELS = ["k_probe_fp64", "k_sdf_field_tiledILi8ELi4ELb0ELb1ELb0", "k_tile_decide", "k_subbox_decide", "k_warp_codes",
           "k_classify_cells", "k_vertex_count", "k_vertex_emitE", "k_vertex_emit_interior", "k_bisect_edges_tiledILb1ELb0",
           "k_tet_emit_fullIj", "k_tet_emitIjLb1", "k_tet_quality_listIj", "k_boundary_sides", "k_boundary_order",
           "k_mark_side_nodesIj", "k_eval_side_nodes", "k_sideset_selectIjLb1", "k_exodus", "k_checksum_tetsIj"]


def main():
    text = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for ln in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = [] if m.group(1) in funcs else funcs.setdefault(m.group(1), [])   # (a second copy of a function is ignored)
            continue
        if cur is not None and re.match(r"\s*/\*[0-9a-f]{4}\*/", ln):
            cur.append(ln.rstrip())
    print("# SASS extract of moosecad_b200/libmoosecad_b200.so (sm_100a, `cuobjdump -sass`)\n")
    print(f"{len(funcs)} device functions.  Opcode counts (static instructions) of the hot kernels:\n")
    print("| kernel | instructions | " + " | ".join(WATCH) + " |")
    print("|---|---|" + "---|" * len(WATCH))
    for k in KERNELS:
        for name, lines in funcs.items():
            if k not in name:
                continue
            c = collections.Counter()
            for ln in lines:
                m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
                if m:
                    c[m.group(1)] += 1
            print(f"| `{name[:60]}` | {len(lines)} | " + " | ".join(str(sum(v for op, v in c.items() if op.startswith(w))) for w in WATCH) + " |")
            break
    for name, lines in funcs.items():
        if "k_warp_codes" in name:
            print("\n## TMA staging in `k_warp_codes` (cp.async.bulk.tensor.3d + mbarrier)\n\n```")
            for i, ln in enumerate(lines):
                if "UTMALDG" in ln or "SYNCS" in ln:
                    for x in lines[max(0, i - 2):i + 2]:
                        print(x)
                    print("        ...")
            print("```")
        if "k_probe_fp64" in name:
            body = [ln for ln in lines if re.search(r"\b(DADD|DMUL|DFMA)\b", ln)]
            print(f"\n## FP64 no-FMA probe `k_probe_fp64`: {sum('DADD' in x for x in body)} DADD, {sum('DMUL' in x for x in body)} DMUL, "
                  f"{sum('DFMA' in x for x in body)} DFMA in the unrolled loop body\n\n```")
            for x in body[:8]:
                print(x)
            print("        ...\n```")
    wide = {name: sum("STG.E.128" in ln for ln in lines) for name, lines in funcs.items()}
    print(f"\n16-byte global stores (`STG.E.128`): {sum(wide.values())} in {sum(1 for v in wide.values() if v)} functions; "
          "the f64 field (`k_sdf_field_tiled`), the tets, the boundary sides and the interior coordinates are written with them.")


if __name__ == "__main__":
    main()
