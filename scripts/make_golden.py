#!/usr/bin/env python
"""Golden fixtures under tests/golden/ (small .npz files).

The reference is Rust with un-vendored crates and cannot be run here, so these vectors are produced by the
ORACLE (oracle/oracle.cpp, libm_mode 1, default assumptions) once it had been pinned to every known answer the
reference's own tests hold (tests/test_oracle_known_answers.py).  They freeze that state: the CPU suite checks
that the oracle still reproduces them bit for bit (no silent drift of the checker), the GPU suite compares the
device path with them without running the oracle.

    python scripts/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle  # noqa: E402
from moosecad_b200 import luascad, scenes  # noqa: E402
import canon  # noqa: E402

CASES = {
    # name: (kind, cells per axis, relative error)
    "cube_res1": ("cube", 1, 0.2),           # BASELINE config 1: examples/cube.lua at --tessellation-resolution 1.0
    "cube_21": ("cube", 21, 0.2),
    "sphere_24": ("sphere", 24, 0.2),
    "xplicit_20": ("xplicit_top", 20, 0.8),
    "csg64_20": ("csg64", 20, 0.2),
}


def build(kind, backend):
    """Returns (volume LObject, {name: boundary LObject}, size of the bounding cube)."""
    if kind.startswith("csg"):
        g = luascad.Geom(backend)
        root = scenes.synthetic_csg(g, int(kind[3:]), seed=int(kind[3:]))
        return root, {"top": scenes.synthetic_csg_top(g, root)}, 1.0
    src = {"cube": scenes.CUBE_LUA, "sphere": scenes.SPHERE_LUA, "xplicit_top": scenes.XPLICIT_WITH_TOP_LUA}[kind]
    out = luascad.eval(src, backend=backend)
    vol = [o.volume() for _, o in sorted(out.items()) if o.volume() is not None][0]
    bnd = {n: o.boundary() for n, o in sorted(out.items()) if o.boundary() is not None}
    return vol, bnd, 15.0 if kind.startswith("xplicit") else 1.0


def oracle_case(kind, n, err):
    pyoracle.set_libm_mode(1)
    pyoracle.reset_assumptions()
    sc = pyoracle.OracleScene()
    vol, bnd, size = build(kind, sc)
    vol.backend.set_parameters_node(vol.node, 0.1, 1.0)
    res = size / n
    m = vol.backend.tessellate(vol.node, res, err)
    assert m.ok()
    sets = {"__boundary__": m.boundary()}
    for name, b in bnd.items():
        b.backend.set_parameters_node(b.node, 0.1, 1.0)
        sets[name] = m.add_sideset(b.backend, b.node)
    c, t, s = canon.canonical(m.vertices(), m.elems(), sets)
    st = m.stats()
    rng = np.random.default_rng(12345)
    bb = np.array(vol.bbox())
    pts = rng.uniform(bb[:3] - 0.1 * size, bb[3:] + 0.1 * size, size=(512, 3))
    vals = vol.backend.eval_points(vol.node, res, pts)
    out = {"coords": c, "tets": t, "points": pts, "values": vals, "resolution": np.float64(res), "error": np.float64(err),
           "dim": np.array(st["dim"], dtype=np.int64), "stats": np.array(st["stats"], dtype=np.int64),
           "ntets_stage": np.array(st["ntets_stage"], dtype=np.int64),
           "n_warped_cut": np.array([st["n_warped"], st["n_cut"]], dtype=np.int64)}
    for name, tri in s.items():
        out["set_" + name] = tri
    return out


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (kind, n, err) in CASES.items():
        d = oracle_case(kind, n, err)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **d)
        print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in d.items()})


if __name__ == "__main__":
    main()
