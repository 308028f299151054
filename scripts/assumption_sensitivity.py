#!/usr/bin/env python
"""Sensitivity of the parity configurations to the recalled details of the
un-vendored crates implicit3d / bbox / nalgebra.  CPU only: the ORACLE is meshed with every named
assumption at its default (what oracle and device implement) and at each alternative reading, and
the resulting meshes are compared (vertex / tet counts, number of canonical tets that differ, max
|delta phi| on the lattice).  An assumption whose alternatives leave every configuration unchanged
cannot make the device path disagree with the real reference on those configurations.

    python scripts/assumption_sensitivity.py [--n 24] > profiles/r02_assumption_sensitivity.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle  # noqa: E402
from moosecad_b200 import luascad, scenes  # noqa: E402
import canon  # noqa: E402


def build(workload, n):
    sc = pyoracle.OracleScene()
    if workload.startswith("csg"):
        g = luascad.Geom(sc)
        root = scenes.synthetic_csg(g, int(workload[3:]), seed=int(workload[3:]))
        root.backend.set_parameters_node(root.node, 0.1, 1.0)
        return root, 1.0 / n, 0.2
    src = {"cube": scenes.CUBE_LUA, "sphere": scenes.SPHERE_LUA, "xplicit": scenes.XPLICIT_LUA,
           "rotated": "c = geom.cube(1, 0.6, 0.4, 0.05):rotate(0.3, 0.5, 0.7)\n"
                      "s = geom.sphere(0.35):translate(0.2, 0.1, 0)\n"
                      "return { v = c:union(s, 0.1):volume() }\n"}[workload]
    out = luascad.eval(src, backend=sc)
    root = next(o.obj for o in out.values() if o.kind == "Volume")
    size = {"cube": 1.0, "sphere": 1.0, "xplicit": 15.0, "rotated": 1.6}[workload]
    return root, size / n, (0.8 if workload == "xplicit" else 0.2)


def mesh_of(workload, n):
    root, res, err = build(workload, n)
    m = root.backend.tessellate(root.node, res, err)
    if not m.ok():
        return None
    return canon.canonical(m.vertices(), m.elems())


def diff(a, b):
    if a is None or b is None:
        return {"failed": True}
    va, ta = a[0], a[1]
    vb, tb = b[0], b[1]
    sa = set(map(bytes, np.ascontiguousarray(va).view(np.uint8).reshape(len(va), -1))) if len(va) else set()
    sb = set(map(bytes, np.ascontiguousarray(vb).view(np.uint8).reshape(len(vb), -1))) if len(vb) else set()
    same = va.shape == vb.shape and ta.shape == tb.shape and np.array_equal(va.view(np.uint64), vb.view(np.uint64)) \
        and np.array_equal(ta, tb)
    return {"identical": bool(same), "vertices": [int(len(va)), int(len(vb))], "tets": [int(len(ta)), int(len(tb))],
            "vertices_only_in_one": int(len(sa ^ sb))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=24)
    ap.add_argument("--workloads", default="cube,sphere,xplicit,rotated,csg16,csg64")
    a = ap.parse_args()
    out = {"what": __doc__.split("\n\n")[0], "lattice_cells_per_axis": a.n, "rows": []}
    for wl in a.workloads.split(","):
        pyoracle.reset_assumptions()
        base = mesh_of(wl, a.n)
        for name, (_, values) in pyoracle.ASSUMPTIONS.items():
            for v, text in values.items():
                if v == 0:
                    continue
                pyoracle.reset_assumptions()
                pyoracle.set_assumption(name, v)
                try:
                    d = diff(base, mesh_of(wl, a.n))
                except Exception as e:  # e.g. an empty bbox under an alternative
                    d = {"error": str(e)}
                d.update(workload=wl, assumption=name, alternative=text)
                out["rows"].append(d)
                print(wl, name, text, d, file=sys.stderr)
    pyoracle.reset_assumptions()
    sens = {}
    for r in out["rows"]:
        sens.setdefault(r["assumption"], []).append(r["workload"]) if not r.get("identical") else None
    out["sensitive_workloads_per_assumption"] = {k: sorted(set(v)) for k, v in sens.items()}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
