#!/usr/bin/env python
"""CPU baseline of BASELINE.md section 3: the oracle (C++ restatement of the reference's serial Rust
path, platform libm) timed at several lattice sizes on the SAME scene, with per-stage times, so that
the dependence of the per-point cost on the resolution is measured instead of assumed.

    python scripts/cpu_baseline_sizes.py --workload csg64 --sizes 40 64 128 [--out profiles/x.json]

One core (the reference has no threads in tessellate3d / mesh / main).  Prints one JSON object.
"""
import argparse
import json
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


def run(workload, n):
    import moosecad_b200 as mc
    from moosecad_b200 import luascad, scenes
    from oracle import pyoracle
    pyoracle.set_libm_mode(0)   # the reference calls the platform libm
    if workload.startswith("csg"):
        nprim = int(workload[3:])
        g = mc.Geom(pyoracle.OracleScene())
        root = scenes.synthetic_csg(g, nprim, seed=nprim, size_scale=0.5 if nprim >= 256 else 1.0)
        res, err, bnds = 1.0 / n, 0.2, []
    else:
        script = {"sphere": scenes.SPHERE_LUA, "xplicit": scenes.XPLICIT_WITH_TOP_LUA, "cube": scenes.CUBE_LUA}[workload]
        out = luascad.eval(script, backend=pyoracle.OracleScene())
        root = [o.volume() for _, o in sorted(out.items()) if o.volume() is not None][0]
        bnds = [o.boundary() for _, o in sorted(out.items()) if o.boundary() is not None]
        bb = root.bbox()
        res, err = (bb[3] - bb[0]) / n, (0.8 if workload == "xplicit" else 0.2)
    root.backend.set_parameters_node(root.node, 0.1, 1.0)
    t0 = time.perf_counter()
    om = root.backend.tessellate(root.node, res, err)
    t_tess = time.perf_counter() - t0
    t0 = time.perf_counter()
    nb = len(om.boundary())
    for b in bnds:
        om.add_sideset(b.backend, b.node)
    t_bnd = time.perf_counter() - t0
    st = om.stats()
    dim = st["dim"]
    pts = (dim[0] + 1) * (dim[1] + 1) * (dim[2] + 1)
    nt = int(st["ntets_stage"][2]) if st["ntets_stage"][2] else len(om.elems())
    nt = len(om.elems())
    return {"n": n, "dim": dim, "lattice_points": pts, "output_tets": nt, "boundary_sides": nb,
            "tessellate_s": t_tess, "boundary_sidesets_s": t_bnd,
            "points_per_s": pts / t_tess, "tets_per_s": nt / t_tess,
            "stage_s": st["times"], "value_calls": st["n_value_calls"], "value_calls_per_point": st["n_value_calls"] / pts}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="csg64")
    ap.add_argument("--sizes", type=int, nargs="+", default=[40, 64])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = [run(a.workload, n) for n in a.sizes]
    doc = {"workload": a.workload, "kind": "port (oracle/oracle.cpp, libm_mode 0 = platform libm)", "cores": 1,
           "cpu": cpu_model(), "host_cores_available": os.cpu_count(), "rows": rows}
    s = json.dumps(doc, indent=1)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")
    print(s)


if __name__ == "__main__":
    main()
