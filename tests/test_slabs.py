"""z-slab sharding.

CPU: partitioning, the exclusive scan of counts and the neighbour exchange over a
world_size-2 gloo group.  GPU: the slab phases on one device against the single-slab mesh
(and, through the oracle, against the reference path)."""
import os
import socket

import numpy as np
import pytest

from moosecad_b200 import slabs

import canon


def test_partition():
    assert slabs.partition(10, 1) == [(0, 10)]
    assert slabs.partition(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert slabs.partition(1024, 8) == [(128 * i, 128 * (i + 1)) for i in range(8)]
    p = slabs.partition(3, 4)
    assert p[-1] == (3, 3) and sum(b - a for a, b in p) == 3
    assert slabs.bases_from_counts([(10, 100), (20, 200), (30, 300)], 2) == (30, 300)
    assert slabs.lattice_dim([-0.5, -0.5, -0.5, 0.5, 0.5, 0.5], 1.0 / 1024) == [1024, 1024, 1024]
    assert slabs.lattice_dim([-7.5] * 3 + [7.5] * 3, 15.0 / 512) == [512, 512, 512]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = slabs.exchange_counts(100 + rank, 1000 * (rank + 1), None, "cpu")
        vb, tb = slabs.bases_from_counts(counts, rank)
        n = 7
        send = torch.arange(n, dtype=torch.int64) + 1000 * rank if rank + 1 < world else None
        recv = torch.zeros(n, dtype=torch.int64) if rank > 0 else None
        slabs.exchange_plane_tables(send, recv, rank, world, None)
        # second round: counts with the top-layer extras, plane table + coordinate slice in one batch
        counts2 = slabs.exchange_counts(100 + rank, 1000 * (rank + 1), None, "cpu", extra=(40 + rank, 2 + rank))
        send_xyz = torch.arange(3 * (2 + rank), dtype=torch.float64) + 0.5 if rank + 1 < world else None
        recv_xyz = torch.zeros(3 * counts2[rank - 1][3], dtype=torch.float64) if rank > 0 else None
        recv2 = torch.zeros(n, dtype=torch.int64) if rank > 0 else None
        slabs.exchange_plane_tables(send, recv2, rank, world, None, send_coords=send_xyz, recv_coords=recv_xyz)
        q.put((rank, counts, vb, tb, None if recv is None else recv.tolist(), counts2,
               None if recv_xyz is None else recv_xyz.tolist(), None if recv2 is None else recv2.tolist()))
    finally:
        dist.destroy_process_group()


def test_count_scan_and_plane_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == [(100, 1000), (101, 2000)]
    assert (res[0][2], res[0][3]) == (0, 0) and (res[1][2], res[1][3]) == (100, 1000)
    assert res[0][4] is None and res[1][4] == list(range(7))   # rank 1 received rank 0's table
    assert res[0][5] == res[1][5] == [(100, 1000, 40, 2), (101, 2000, 41, 3)]
    assert res[0][6] is None and res[1][6] == [0.5 + i for i in range(6)]   # ... and its coordinate slice
    assert res[1][7] == list(range(7))


@pytest.mark.gpu
@pytest.mark.parametrize("n_slabs", [2, 3, 5])
@pytest.mark.parametrize("scene", ["sphere", "csg"])
def test_slabs_match_single_slab(scene, n_slabs, ctx):
    import moosecad_b200 as mc
    from moosecad_b200 import luascad, scenes
    if scene == "sphere":
        obj, res = luascad.eval(scenes.SPHERE_LUA)["sphere"].obj, 1.0 / 37
    else:
        g = mc.Geom()
        obj, res = scenes.synthetic_csg(g, 12, seed=5), 1.0 / 45
    obj.backend.set_parameters(0.1, 1.0)
    whole = mc.IsosurfaceStuffing(obj, res, 0.2, ctx=ctx).tessellate()
    coords, tets, stats = slabs.tessellate_slabs_one_device(ctx, obj, res, 0.2, n_slabs, return_stats=True)
    # check_for_inverted_tets over the slabs = over the whole mesh: every tet is evaluated exactly
    # once, the ones touching a shared plane with the coordinates received from the slab below
    ws = whole.stats()
    for key in ("ntets_stage", "stats", "aspect_ratio", "inverted"):
        assert stats[key] == ws[key], key
    a = canon.canonical(whole.vertices(), whole.elems())
    b = canon.canonical(coords, tets)
    canon.assert_same_mesh(a, b)
    # global ids: every vertex referenced, none duplicated
    assert np.unique(tets).size == len(coords)
    assert len(np.unique(coords, axis=0)) == len(coords)


@pytest.mark.gpu
def test_vertex_ids_beyond_32_bits(ctx):
    """Global vertex ids >= 2^32 (a slab high up in a very large lattice) switch the tet emission, the quality
    list, the checksums and the boundary to 64-bit indices: same mesh, ids shifted."""
    import moosecad_b200 as mc
    g = mc.Geom()
    from moosecad_b200 import scenes
    obj, res = scenes.synthetic_csg(g, 12, seed=5), 1.0 / 41
    obj.backend.set_parameters(0.1, 1.0)
    base = (1 << 32) + 12345
    c0, t0, s0 = slabs.tessellate_slabs_one_device(ctx, obj, res, 0.2, 2, return_stats=True)
    c1, t1, s1 = slabs.tessellate_slabs_one_device(ctx, obj, res, 0.2, 2, return_stats=True, id_offset=base)
    assert np.array_equal(c0.view(np.uint64), c1.view(np.uint64))
    assert t1.min() >= base and np.array_equal(t1 - np.uint64(base), t0)
    assert s0 == s1


def test_partition_balanced():
    c = [0.02] * 20 + [1.0] * 60 + [0.02] * 20
    for w in (2, 4, 8):
        p = slabs.partition_balanced(c, w)
        assert p[0][0] == 0 and p[-1][1] == len(c) and all(a[1] == b[0] for a, b in zip(p, p[1:]))
        loads = [sum(c[a:b]) for a, b in p]
        assert max(loads) <= 1.15 * sum(c) / w
        assert all(b > a for a, b in p)
    assert slabs.partition_balanced([0.0] * 10, 2) == [(0, 5), (5, 10)]
    assert slabs.partition_balanced([1.0] * 3, 1) == [(0, 3)]


def test_costs_from_classes_arithmetic():
    """Host side of the sharded load estimate: per-cell-layer cost = weights . class table row / 16."""
    w = slabs.class_weights()
    assert len(w) == 7 and all(x >= 0.0 for x in w) and w[5] > 0.0
    table = np.zeros((3, 7), dtype=np.uint64)
    table[0, 0], table[0, 5] = 10, 50000
    table[1, 1] = 7
    table[2, 3] = 100
    costs = slabs.costs_from_classes(table, 40)
    assert len(costs) == 40
    assert costs[0] == costs[15] == (w[0] * 10 + w[5] * 50000) / 16.0
    assert costs[16] == costs[31] == w[1] * 7 / 16.0
    assert costs[32] == costs[39] == w[3] * 100 / 16.0


@pytest.mark.gpu
def test_sharded_load_estimate_matches_whole(ctx):
    """mc_tile_layer_classes over ranges of tile layers sums to the table of one full pass, and the
    costs derived from it equal mc_estimate_layer_costs."""
    import moosecad_b200 as mc
    from moosecad_b200 import scenes
    g = mc.Geom()
    root = scenes.synthetic_csg(g, 16)
    res = 1.0 / 96
    whole = slabs.tile_layer_classes(ctx, root, res)
    n = whole.shape[0]
    assert whole[:, :5].sum() == n * 7 * 7            # 97 points per axis -> 7 tiles
    assert whole[:, 1:5].sum() > 0 and whole[:, 0].sum() > 0
    for world in (2, 3, n):
        acc = np.zeros_like(whole)
        for k0, k1 in slabs.partition(n, world):
            part = slabs.tile_layer_classes(ctx, root, res, k0, k1)
            assert not part[:k0].any() and not part[k1:].any()
            acc += part
        assert np.array_equal(acc, whole)
    dim = slabs.lattice_dim(root.bbox(), res)
    direct = slabs.layer_costs(ctx, root, res)
    assert np.allclose(direct, slabs.costs_from_classes(whole, dim[2]), rtol=1e-12, atol=0.0)


def test_combine_stats():
    """Per-slab statistics add up (counters) / combine by hull (aspect-ratio range)."""
    a = {"ntets_stage": [10, 8, 9], "stats": [1, 0, 2, 0, 0, 3, 0], "aspect_ratio": [0.5, 0.9], "inverted": 1}
    b = {"ntets_stage": [5, 5, 5], "stats": [0, 1, 0, 0, 4, 0, 0], "aspect_ratio": [0.4, 0.8], "inverted": 0}
    c = slabs.combine_stats([a, b])
    assert c == {"ntets_stage": [15, 13, 14], "stats": [1, 1, 2, 0, 4, 3, 0], "aspect_ratio": [0.4, 0.9], "inverted": 1}
    # the reference starts the range at (one, zero) (tessellate3d/src/lib.rs:404): an empty slab is neutral
    empty = {"ntets_stage": [0, 0, 0], "stats": [0] * 7, "aspect_ratio": [1.0, 0.0], "inverted": 0}
    assert slabs.combine_stats([a, empty]) == slabs.combine_stats([a])


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["xplicit_top", "cube"])
@pytest.mark.parametrize("n_slabs", [2, 3])
def test_tessellate_multi_single_process(n_slabs, scene, ctx):
    """mc_tessellate_multi (one process, one host thread per context, peer copies of the plane
    table): several contexts on the ONE device of the test box, against the single-slab mesh —
    mesh, statistics, device checksums, boundary sides and a side-set (slab-aware)."""
    import moosecad_b200 as mc
    from moosecad_b200 import _lib, luascad, scenes
    from moosecad_b200.tessellate3d import mesh_checksum
    import ctypes as C
    out = luascad.eval(scenes.XPLICIT_WITH_TOP_LUA if scene == "xplicit_top" else scenes.CUBE_LUA)
    obj = [o.volume() for _, o in sorted(out.items()) if o.volume() is not None][0]
    top = [o.boundary() for n, o in sorted(out.items()) if o.boundary() is not None and n == "top"][0]
    obj.backend.set_parameters(0.1, 1.0)
    res = 15.0 / 44 if scene == "xplicit_top" else 1.0 / 21
    whole = mc.IsosurfaceStuffing(obj, res, 0.2, ctx=ctx).tessellate()
    ctxs = [mc.Context(0) for _ in range(n_slabs)]
    handles = slabs.tessellate_multi(ctxs, obj, res, 0.2)
    lib = _lib.lib()
    try:
        coords, tets, stats = slabs.gather_slab_meshes(handles)
        ws = whole.stats()
        for key in ("ntets_stage", "stats", "aspect_ratio", "inverted"):
            assert stats[key] == ws[key], key
        canon.assert_same_mesh(canon.canonical(whole.vertices(), whole.elems()), canon.canonical(coords, tets))
        cs = [mesh_checksum(h) for h in handles]
        wcs = whole.checksum()
        assert sum(c["vertices"] for c in cs) % (1 << 64) == wcs["vertices"]
        assert sum(c["tets"] for c in cs) % (1 << 64) == wcs["tets"]
        assert sum(c["tets_skipped"] for c in cs) == 0
        # Mesh::boundary and the side-set loop on slabs: global (elem, side) pairs, no exchange
        sides, members = [], []
        for h in handles:
            n = C.c_uint64()
            _lib.check(lib.mc_mesh_boundary(h, C.byref(n)))
            if n.value:
                sides.append(np.ctypeslib.as_array(lib.mc_mesh_boundary_sides(h), shape=(n.value, 2)).copy())
            idx = _lib.check(lib.mc_mesh_add_sideset(h, top.backend.handle, top.node))
            k = int(lib.mc_mesh_sideset_len(h, idx))
            if k:
                members.append(np.ctypeslib.as_array(lib.mc_mesh_sideset(h, idx), shape=(k, 2)).copy())
        sides = np.concatenate(sides)
        members = np.concatenate(members) if members else np.zeros((0, 2), dtype=np.uint64)
        wtop = whole.add_boundary_side_set("top", top)
        a = canon.canonical(whole.vertices(), whole.elems(), {"s": sorted(whole.boundary().sides()), "top": sorted(wtop.sides())})[2]
        b = canon.canonical(coords, tets, {"s": sides, "top": members})[2]
        assert np.array_equal(a["s"], b["s"]) and np.array_equal(a["top"], b["top"])
        if scene == "cube":
            assert len(a["top"]) > 0
    finally:
        for h in handles:
            lib.mc_mesh_free(h)


def test_measured_load_refinement_converges():
    """refine_partition: slabs cut from a wrong cost estimate, rebalanced with the times the ranks measure."""
    rng = np.random.default_rng(1)
    n, world = 1024, 8
    true = 0.2 + np.abs(np.sin(np.arange(n) / 90.0)) + 0.5 * rng.random(n)      # what a layer really costs
    est = list(1.0 + 0.0 * true)                                                  # a flat (wrong) estimate
    ent = slabs._Partition(est, slabs.partition_balanced(est, world))
    def times(sl):
        return [int(1e3 * true[a:b].sum()) for a, b in sl]
    t0 = times(ent.slabs)
    for _ in range(slabs.FEEDBACK_ROUNDS + 2):
        slabs.refine_partition(ent, times(ent.slabs))
    t1 = times(ent.slabs)
    assert ent.rounds <= slabs.FEEDBACK_ROUNDS
    assert ent.slabs[0][0] == 0 and ent.slabs[-1][1] == n and all(a[1] == b[0] for a, b in zip(ent.slabs, ent.slabs[1:]))
    assert max(t1) / (sum(t1) / world) < 1.03 < max(t0) / (sum(t0) / world)
    # a missing time (a rank that has not reported yet) changes nothing
    before = list(ent.slabs)
    ent.rounds = 0
    assert not slabs.refine_partition(ent, [0] + times(ent.slabs)[1:]) and ent.slabs == before


def test_refinement_keeps_the_best_partition_seen():
    """A refinement round that makes the slowest rank slower is undone: the best partition measured so far comes
    back and the refinement stops (the cost model inside a slab is only a rescaled estimate)."""
    n, world = 256, 4
    ent = slabs._Partition([1.0] * n, slabs.partition_balanced([1.0] * n, world))
    first = list(ent.slabs)
    assert slabs.refine_partition(ent, [900, 1000, 1100, 1000]) and ent.slabs != first and ent.rounds == 1
    assert ent.best == (1100, first)
    # the new slabs turn out worse (slowest rank 1300 > 1100): back to the first partition, no further rounds
    assert slabs.refine_partition(ent, [700, 1300, 1000, 1000]) and ent.slabs == first
    assert ent.rounds == slabs.FEEDBACK_ROUNDS
    assert not slabs.refine_partition(ent, [900, 1000, 1100, 1000]) and ent.slabs == first
    # ... while a round that helps is kept and becomes the new best
    ent2 = slabs._Partition([1.0] * n, slabs.partition_balanced([1.0] * n, world))
    slabs.refine_partition(ent2, [900, 1000, 1100, 1000])
    second = list(ent2.slabs)
    slabs.refine_partition(ent2, [1000, 1010, 1030, 1000])
    assert ent2.best == (1030, second)
