#!/usr/bin/env python
"""One meshing step of a bench workload between cudaProfilerStart / Stop, after two warm steps:

    ncu --profile-from-start off --set full --clock-control none --import-source on \\
        -k regex:'k_warp_codes|k_classify' -o gpurun_out/x python scripts/profile_step.py csg64_1024

(no timing is taken here: a number measured under a profiler is never a bench value)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import moosecad_b200 as mc  # noqa: E402
from moosecad_b200 import _lib  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "csg64_1024"
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    with_boundary = not (len(sys.argv) > 3 and sys.argv[3] == "noboundary")
    lib = _lib.lib()
    ctx = mc.Context(0)
    g, root, res, n, err, bnds = bench.build_scene(mc.Geom, name)
    g.backend.set_parameters(0.1, 1.0)

    def step():
        opts = _lib.mc_options(C.sizeof(_lib.mc_options), flags, 0, 0)
        out = C.c_void_p()
        rc = lib.mc_tessellate(ctx.handle, root.backend.handle, root.node, res, err, C.byref(opts), C.byref(out))
        if rc not in (_lib.MC_OK, _lib.MC_ERR_DIVERGED):
            _lib.check(rc)
        if with_boundary:
            ns = C.c_uint64()
            _lib.check(lib.mc_mesh_boundary(out, C.byref(ns)))
            for _, b in bnds:
                _lib.check(lib.mc_mesh_add_sideset(out, b.backend.handle, b.node))
        lib.mc_mesh_free(out)

    step()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
