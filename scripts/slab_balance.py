"""Per-slab device time of the z-slab partitions, measured on ONE GPU (slab after slab).

    python scripts/slab_balance.py [--workload csg64_1024] [--worlds 2,4,8]

For every world size prints, for the equal and the cost-balanced partition, each slab's wall time
(begin -> emit_vertices -> emit_tets, synchronised) and the time of the cost estimate itself.
The slowest slab bounds the multi-GPU step; the spread shows what the estimate leaves on the table.
"""
import argparse
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from moosecad_b200 import _lib, slabs  # noqa: E402
import moosecad_b200 as mc  # noqa: E402


def run_slab(lib, ctx, g, root, res, z0, z1, lower, bases):
    slab = _lib.mc_slab(z0, z1)
    opts = _lib.mc_options(C.sizeof(_lib.mc_options), 0, 0, 0)
    out = C.c_void_p()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(lib.mc_tessellate_begin(ctx.handle, g.backend.handle, root.node, res, bench.REL_ERR, C.byref(slab),
                                       C.byref(opts), C.byref(out)))
    c = (C.c_uint64 * 6)()
    _lib.check(lib.mc_mesh_counts(out, c))
    # global ids: the table of the slab below refers to ITS vertex range
    _lib.check(lib.mc_tessellate_emit_vertices(out, bases[0]))
    rc = lib.mc_tessellate_emit_tets(out, bases[1], lower)
    bases[0] += c[0]
    bases[1] += c[1]
    if rc not in (_lib.MC_OK, _lib.MC_ERR_DIVERGED):
        _lib.check(rc)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    st = (C.c_double * 9)()
    lib.mc_mesh_stage_ms(out, st)
    return out, ms, list(st)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="csg64_1024")
    ap.add_argument("--worlds", default="2,4,8")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    lib = _lib.lib()
    with torch.cuda.stream(stream):
        ctx = mc.Context(0, stream.cuda_stream)
        g, root, res, n = bench.build_scene(mc.Geom, args.workload)
        g.backend.set_parameters(0.1, 1.0)
        dim = slabs.lattice_dim(root.bbox(), res)
        for _ in range(2):
            slabs.layer_costs(ctx, root, res)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        costs = slabs.layer_costs(ctx, root, res).tolist()
        torch.cuda.synchronize()
        print(f"layer cost estimate: {(time.perf_counter() - t0) * 1e3:.2f} ms")
        import numpy as np
        classes = slabs.tile_layer_classes(ctx, root, res).astype(np.float64)   # (tile layers, 7)
        rows, obs = [], []

        def slab_classes(z0, z1):
            # cell layers [z0, z1) -> fraction of every 16-layer tile layer
            acc = np.zeros(classes.shape[1])
            for k in range(classes.shape[0]):
                lo, hi = max(z0, 16 * k), min(z1, 16 * k + 16)
                if hi > lo:
                    acc += classes[k] * (hi - lo) / 16.0
            return acc

        for world in [int(w) for w in args.worlds.split(",")]:
            for name, part in (("equal", slabs.partition(dim[2], world)),
                               ("balanced", slabs.partition_balanced(costs, world))):
                for rep in range(2):           # first repetition warms the pool
                    times, stages = [], []
                    prev = None
                    bases = [0, 0]
                    for (z0, z1) in part:
                        lower = None
                        if prev is not None:
                            table, cnt = C.c_void_p(), C.c_uint64()
                            _lib.check(lib.mc_mesh_upper_plane_table(prev, C.byref(table), C.byref(cnt)))
                            lower = table if cnt.value else None
                        h, ms, st = run_slab(lib, ctx, g, root, res, z0, z1, lower, bases)
                        if prev is not None:
                            lib.mc_mesh_free(prev)
                        prev = h
                        times.append(ms)
                        stages.append(st)
                        if rep == 1:
                            rows.append(np.concatenate([slab_classes(z0, z1), [1.0]]))
                            obs.append(ms)
                    lib.mc_mesh_free(prev)
                print(f"world {world} {name:8s} slabs {part}")
                print("   ms/slab  " + " ".join(f"{t:6.1f}" for t in times) + f"   max {max(times):.1f} mean {sum(times) / len(times):.1f}")
                print("   sdf      " + " ".join(f"{s[0]:6.1f}" for s in stages))
                print("   rest     " + " ".join(f"{sum(s[1:]):6.1f}" for s in stages))
        A, y = np.array(rows), np.array(obs)
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez("gpurun_out/slab_balance_rows.npz", A=A, y=y)
        w, *_ = np.linalg.lstsq(A, y, rcond=None)
        print("least squares: ms per [unproven, neg skipped, neg shell, pos skipped, pos shell] tile, per unit of "
              "[unproven, shell] work, per-slab constant")
        print("   ", " ".join(f"{x:.3e}" for x in w))
        print(f"    residual rms {np.sqrt(np.mean((A @ w - y) ** 2)):.2f} ms over {len(y)} slabs")


if __name__ == "__main__":
    main()
