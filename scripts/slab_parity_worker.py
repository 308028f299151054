"""torchrun worker: mesh a scene as z-slabs over all ranks (NCCL), gather the slab meshes on
rank 0 and compare them, after canonical sorting, with the single-GPU mesh of the same scene.
Used by tests/test_multi_gpu.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/slab_parity_worker.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import moosecad_b200 as mc
from moosecad_b200 import _lib, luascad, scenes, slabs
import canon


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    host_group = dist.new_group(backend="gloo")
    lib = _lib.lib()
    stream = torch.cuda.Stream(device=dev)
    ok = True
    with torch.cuda.stream(stream):
        ctx = mc.Context(local, stream.cuda_stream)
        cases = []
        cases.append(("sphere", luascad.eval(scenes.SPHERE_LUA)["sphere"].obj, 1.0 / 61))
        g = mc.Geom()
        cases.append(("csg", scenes.synthetic_csg(g, 20, seed=3), 1.0 / 90))
        for name, obj, res in cases:
            obj.backend.set_parameters(0.1, 1.0)
            # the sphere takes the NCCL-only path, the CSG scene sends its host-valued bookkeeping
            # (class table, counts) over the gloo host group like bench.py does
            sm = slabs.SlabMesher(ctx, obj, res, 0.2, rank, world, None, dev,
                                  host_group=host_group if name == "csg" else None)
            h = sm.run()
            cnt = (C.c_uint64 * 6)()
            lib.mc_mesh_counts(h, cnt)
            coords = np.empty((cnt[0], 3), dtype=np.float64)
            tets = np.empty((cnt[1], 4), dtype=np.uint64)
            _lib.check(lib.mc_mesh_copy_to_host(h, coords.ctypes.data_as(C.c_void_p), tets.ctypes.data_as(C.c_void_p), 8))
            st = slabs.mesh_stats(h)
            # Mesh::boundary on the slab (global (elem, side) pairs) and the device checksums
            from moosecad_b200.tessellate3d import mesh_checksum
            ns = C.c_uint64()
            _lib.check(lib.mc_mesh_boundary(h, C.byref(ns)))
            sides = np.ctypeslib.as_array(lib.mc_mesh_boundary_sides(h), shape=(ns.value, 2)).copy() if ns.value else np.zeros((0, 2), dtype=np.uint64)
            cs = mesh_checksum(h)
            lib.mc_mesh_free(h)
            gathered = [None] * world if rank == 0 else None
            dist.gather_object((coords, tets, sm.vertex_base, sm.tet_base, st, sides, cs), gathered, dst=0)
            if rank == 0:
                gathered.sort(key=lambda t: t[2])
                allc = np.concatenate([t[0] for t in gathered])
                allt = np.concatenate([t[1] for t in gathered])
                whole = mc.IsosurfaceStuffing(obj, res, 0.2, ctx=ctx).tessellate()
                try:
                    canon.assert_same_mesh(canon.canonical(whole.vertices(), whole.elems()), canon.canonical(allc, allt))
                    assert np.unique(allt).size == len(allc)
                    ws, ss = whole.stats(), slabs.combine_stats([t[4] for t in gathered])
                    for key in ("ntets_stage", "stats", "aspect_ratio", "inverted"):
                        assert ss[key] == ws[key], (key, ss[key], ws[key])
                    alls = np.concatenate([t[5] for t in gathered])
                    a = canon.canonical(whole.vertices(), whole.elems(), {"s": sorted(whole.boundary().sides())})[2]["s"]
                    b = canon.canonical(allc, allt, {"s": alls})[2]["s"]
                    assert np.array_equal(a, b), "boundary sides differ"
                    wcs = whole.checksum()
                    assert sum(t[6]["vertices"] for t in gathered) % (1 << 64) == wcs["vertices"]
                    assert sum(t[6]["tets"] for t in gathered) % (1 << 64) == wcs["tets"]
                    assert sum(t[6]["tets_skipped"] for t in gathered) == 0
                    print(f"SLAB_PARITY_OK {name} world={world} verts={len(allc)} tets={len(allt)}", flush=True)
                except AssertionError as e:
                    ok = False
                    print(f"SLAB_PARITY_FAIL {name}: {e}", flush=True)
        del ctx
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
