"""Exploratory timing of the large BASELINE configs on one GPU (not a benchmark)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import moosecad_b200 as mc
from moosecad_b200 import luascad, scenes, _lib

ctx = mc.Context(0)
print("fp64 no-fma probe: %.2f TFLOP/s" % (ctx.probe_fp64_nofma() / 1e12))
which = sys.argv[1:] or ["sphere256", "xplicit512", "csg64_512", "csg64_1024"]
for w in which:
    err = 0.2
    if w == "sphere256":
        obj = luascad.eval(scenes.SPHERE_LUA)["sphere"].obj; res = 1.0 / 256
    elif w == "xplicit512":
        obj = luascad.eval(scenes.XPLICIT_LUA)["xplicit"].obj; res = 15.0 / 512; err = 0.8
    elif w.startswith("csg"):
        npr, n = w[3:].split("_")
        g = mc.Geom(); obj = scenes.synthetic_csg(g, int(npr)); res = 1.0 / int(n)
    obj.backend.set_parameters(0.1, 1.0)
    info = obj.backend.program_info(obj.node, res)
    for rep in range(2):
        t0 = time.time()
        mesh = mc.IsosurfaceStuffing(obj, res, err, ctx=ctx, allow_diverged=True).tessellate()
        dt = time.time() - t0
        st = mesh.stats()
        if rep == 1:
            t1 = time.time(); nb = len(mesh.boundary()) if "1024" not in w else -1; tb = time.time() - t1
            print(w, "dim", st["dim"], "verts", st["n_vertices"], "tets", st["n_tets"], "cut", st["n_cut"], "warped", st["n_warped"],
                  "wall %.3fs" % dt, "stage_ms", [round(x, 3) for x in st["stage_ms"]], "boundary", nb, "%.3fs" % tb,
                  "prog", info, "ws GB %.1f" % (ctx.workspace_bytes / 1e9), "ar", st["aspect_ratio"], "inv", st["inverted"], "prune", st["prune"], "tiles", st["tiles"], flush=True)
        del mesh
