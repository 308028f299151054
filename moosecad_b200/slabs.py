"""z-slab sharding of one meshing job over the ranks of a torch.distributed group
(SURVEY.md §8e): one process per GPU, no data-path collective except

  * an all-gather of the per-slab (vertex, tet) counts -> global index bases, and
  * one neighbour exchange of the id table of the shared lattice plane (rank r -> r + 1).

Every rank evaluates its own slab plus a two-plane halo (the SDF is a pure function of
position, so the halo is recomputed instead of exchanged).  torch is used for device
buffers, streams and the process group only.
"""
import ctypes as C
import math

from . import _lib


def lattice_dim(bbox, resolution):
    """IsosurfaceStuffing::new (tessellate3d/src/lib.rs:43-47): ceil(bbox.dim / resolution)."""
    return [int(math.ceil((bbox[3 + k] - bbox[k]) / resolution)) for k in range(3)]


def partition(nz, world_size):
    """Contiguous z ranges [z0, z1) of cell layers, as even as possible (earlier ranks get the
    remainder).  Ranks beyond nz get empty ranges."""
    base, rem = divmod(int(nz), int(world_size))
    out, z = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append((z, z + n))
        z += n
    return out


def partition_balanced(costs, world_size):
    """Contiguous z ranges with (nearly) equal summed cost; every rank gets at least one layer
    when there are enough layers."""
    n = len(costs)
    if world_size <= 1 or n <= world_size:
        return partition(n, world_size)
    total = float(sum(costs))
    if not total > 0.0:
        return partition(n, world_size)
    out, z, acc = [], 0, 0.0
    for r in range(world_size):
        remaining_ranks = world_size - r
        if remaining_ranks == 1:
            out.append((z, n))
            break
        target = (total - acc) / remaining_ranks
        z1, s = z, 0.0
        # take layers until the target is reached, leaving at least one layer per remaining rank
        while z1 < n - (remaining_ranks - 1) and (z1 == z or s + 0.5 * costs[z1] <= target):
            s += costs[z1]
            z1 += 1
        out.append((z, z1))
        acc += s
        z = z1
    return out


def layer_costs(ctx, obj, resolution):
    """Per-cell-layer load estimate (C ABI mc_estimate_layer_costs)."""
    import numpy as np
    dim = lattice_dim(obj.bbox(), float(resolution))
    cost = np.zeros(dim[2], dtype=np.float64)
    if dim[2]:
        _lib.check(_lib.lib().mc_estimate_layer_costs(ctx.handle, obj.backend.handle, obj.node, float(resolution),
                                                      cost.ctypes.data_as(_lib._PD), dim[2]))
    return cost


def combine_stats(per_slab):
    """Statistics of the whole mesh from the per-slab ones (each a dict with "ntets_stage", "stats",
    "aspect_ratio", "inverted" as Mesh.stats() returns them): counters add up, the aspect-ratio
    range is the hull."""
    out = {"ntets_stage": [0, 0, 0], "stats": [0] * 7, "aspect_ratio": [1.0, 0.0], "inverted": 0}
    for st in per_slab:
        out["ntets_stage"] = [a + b for a, b in zip(out["ntets_stage"], st["ntets_stage"])]
        out["stats"] = [a + b for a, b in zip(out["stats"], st["stats"])]
        out["aspect_ratio"] = [min(out["aspect_ratio"][0], st["aspect_ratio"][0]),
                               max(out["aspect_ratio"][1], st["aspect_ratio"][1])]
        out["inverted"] += st["inverted"]
    return out


def mesh_stats(handle):
    """The additive part of Mesh.stats() for a raw mc_mesh handle."""
    lib = _lib.lib()
    dim, stage, st = (C.c_uint64 * 3)(), (C.c_uint64 * 3)(), (C.c_uint64 * 7)()
    ar, inv = (C.c_double * 2)(), C.c_uint64()
    _lib.check(lib.mc_mesh_stats(handle, dim, stage, st, ar, C.byref(inv)))
    return {"dim": list(dim), "ntets_stage": list(stage), "stats": list(st), "aspect_ratio": list(ar), "inverted": inv.value}


def bases_from_counts(all_counts, rank):
    """Exclusive scan over ranks of (n_vertices, n_tets)."""
    vb = sum(int(c[0]) for c in all_counts[:rank])
    tb = sum(int(c[1]) for c in all_counts[:rank])
    return vb, tb


_side_streams = {}


def _side_stream(device):
    """A second CUDA stream per device for the small host-synchronous collectives, so that they do
    not wait for (or hold up) the kernels queued on the meshing stream."""
    import torch
    key = str(device)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def exchange_counts(n_vertices, n_tets, group=None, device="cpu", side_stream=False, extra=()):
    """all-gather of a few int64 per rank; returns a list of (n_vertices, n_tets, *extra) per rank.
    With side_stream=True (CUDA devices) the collective runs on its own stream: the counts are
    host values, nothing on the meshing stream has to finish first."""
    import contextlib
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    on_cuda = str(device).startswith("cuda")
    cm = torch.cuda.stream(_side_stream(device)) if (side_stream and on_cuda) else contextlib.nullcontext()
    with cm:
        vals = [int(n_vertices), int(n_tets)] + [int(x) for x in extra]
        mine = torch.tensor(vals, dtype=torch.int64, device=device)
        out = [torch.zeros(len(vals), dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(out, mine, group=group)
        return [tuple(int(x) for x in t.tolist()) for t in out]


def exchange_plane_tables(send_table, recv_table, rank, world, group=None, active=None, send_coords=None,
                          recv_coords=None):
    """Rank r sends its upper-plane id table to r + 1 and receives r - 1's.  `send_table` /
    `recv_table` are int64 tensors (or None at the ends).  `active[r]` tells whether rank r
    owns any cell layer; empty ranks are skipped by relaying nothing.  `send_coords` /
    `recv_coords` (float64, optional, possibly empty) carry the coordinates of the sender's top
    tile layer in the same batch."""
    import torch.distributed as dist
    ops = []
    if send_table is not None and rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, send_table, rank + 1, group))
        if send_coords is not None and send_coords.numel():
            ops.append(dist.P2POp(dist.isend, send_coords, rank + 1, group))
    if recv_table is not None and rank > 0:
        ops.append(dist.P2POp(dist.irecv, recv_table, rank - 1, group))
        if recv_coords is not None and recv_coords.numel():
            ops.append(dist.P2POp(dist.irecv, recv_coords, rank - 1, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def tile_layer_classes(ctx, obj, resolution, layer_begin=0, layer_end=None):
    """Tile counts (and work sums) per tile layer and cost class (C ABI mc_tile_layer_classes):
    (n_tile_layers, 7); only the rows [layer_begin, layer_end) are computed, the rest are zero."""
    import numpy as np
    dim = lattice_dim(obj.bbox(), float(resolution))
    n = (dim[2] + 1 + 15) // 16
    counts = np.zeros((n, _lib.MC_TILE_CLASS_STRIDE), dtype=np.uint64)
    end = n if layer_end is None else int(layer_end)
    _lib.check(_lib.lib().mc_tile_layer_classes(ctx.handle, obj.backend.handle, obj.node, float(resolution),
                                                int(layer_begin), end, counts.ctypes.data_as(C.POINTER(C.c_uint64)), n))
    return counts


def class_weights():
    w = (C.c_double * _lib.MC_TILE_CLASS_STRIDE)()
    _lib.lib().mc_slab_class_weights(w)
    return list(w)


def costs_from_classes(counts, nz):
    """Relative cost of every CELL layer from the per-tile-layer class table (the same arithmetic
    as mc_estimate_layer_costs)."""
    w = class_weights()
    per_tile_layer = [sum(w[q] * float(row[q]) for q in range(len(w))) for row in counts]
    return [per_tile_layer[z // 16] / 16.0 for z in range(nz)]


def layer_costs_sharded(ctx, obj, resolution, rank, world, group=None, device=None, host_group=None):
    """Load estimate with the interval pass itself sharded: every rank classifies 1/world of the tile
    layers, one all-reduce (SUM) of the small class table makes it whole on every rank."""
    import torch
    import torch.distributed as dist
    dim = lattice_dim(obj.bbox(), float(resolution))
    n = (dim[2] + 1 + 15) // 16
    k0, k1 = partition(n, world)[rank]
    counts = tile_layer_classes(ctx, obj, resolution, k0, k1)
    t = torch.from_numpy(counts.astype("int64"))
    if host_group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=host_group)   # host values over a host (gloo) group
    else:
        if device is not None and str(device) != "cpu":
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return costs_from_classes(t.cpu().numpy(), dim[2])


_partition_cache = {}
# Rounds of measured-load refinement (0 disables it).  A round uses the device times of the previous run of the
# scene, shared by a tiny all-gather BEFORE the run starts, so the slabs change at the start of runs 2..5 and are
# final from then on (inside a 5-step warm-up); the best partition seen wins if a round makes things worse.
FEEDBACK_ROUNDS = int(__import__("os").environ.get("MC_SLAB_FEEDBACK", "4"))


class _Partition:
    """Cache entry of one (scene, resolution, world): the slabs in use, the per-layer cost estimate they
    came from, this rank's device time of the last run with them (reported by the caller) and the
    number of measured-load refinements applied so far."""

    def __init__(self, costs, slabs):
        self.costs, self.slabs = list(costs), slabs
        self.my_last_us = 0       # device time of this rank's last run with `slabs` (microseconds, 0 = unknown)
        self.rounds = 0
        self.best = None          # (slowest rank's time, slabs) of the best partition measured so far


def cached_partition(ctx, obj, resolution, rank, world, group=None, device=None, host_group=None):
    """Balanced z-slab partition of (scene, resolution, world), computed once per process: the load
    estimate is a deterministic function of the scene's device program, so the key is its content
    hash (a scene lowered again from the same script hits the cache).  Every rank must call it in
    the same order (the first call for a key runs one collective)."""
    res = float(resolution)
    key = (obj.backend.program_hash(obj.node, res), res, int(world))
    ent = _partition_cache.get(key)
    if ent is None:
        costs = layer_costs_sharded(ctx, obj, res, rank, world, group, device, host_group)
        ent = _Partition(costs, partition_balanced(costs, world))
        _partition_cache[key] = ent
    return ent


def refine_partition(ent, times_us):
    """Measured-load refinement: the cost estimate of every slab is rescaled so that it sums to the
    device time its rank just measured, and the lattice is cut again.  Deterministic in its inputs
    (every rank holds the same times after the all-gather), applied a few times only
    (FEEDBACK_ROUNDS), skipped when a time is missing or the slabs are already within 1.5 %."""
    if ent.rounds >= FEEDBACK_ROUNDS or any(t <= 0 for t in times_us):
        return False
    mean = sum(times_us) / len(times_us)
    best = getattr(ent, "best", None)
    if best is not None and max(times_us) > best[0]:
        # the last round made the slowest rank slower: back to the best partition seen, and stop
        ent.slabs = best[1]
        ent.rounds = FEEDBACK_ROUNDS
        return True
    ent.best = (max(times_us), list(ent.slabs))
    if max(times_us) <= 1.015 * mean:
        ent.rounds = FEEDBACK_ROUNDS
        return False
    costs = list(ent.costs)
    for (z0, z1), t in zip(ent.slabs, times_us):
        est = sum(costs[z0:z1])
        if est > 0.0:
            f = t / est
            for z in range(z0, z1):
                costs[z] *= f
    ent.costs = costs
    ent.slabs = partition_balanced(costs, len(times_us))
    ent.rounds += 1
    return True


def _ctx_stream(ctx, device):
    """The torch view of the CUDA stream a Context launches on."""
    import torch
    if getattr(ctx, "stream", 0):
        return torch.cuda.ExternalStream(ctx.stream, device=device)
    return torch.cuda.default_stream(device)


class SlabMesher:
    """Meshes this rank's slab of the global lattice.  Usage (every rank):

        sm = SlabMesher(ctx, obj, res, err, rank, world, group, device)
        mesh_handle = sm.run()        # device-resident slab mesh with GLOBAL vertex ids
    """

    def __init__(self, ctx, obj, resolution, relative_error, rank, world, group=None, device=None, flags=0,
                 balanced=True, host_group=None):
        """group: the device (NCCL) group the plane tables travel over.  host_group: optional host
        (gloo) group for the two tiny host-valued exchanges (class table, counts); without it they
        use `group` too, which on NCCL queues them behind the kernels already launched."""
        self.ctx, self.obj = ctx, obj
        self.res, self.err = float(resolution), float(relative_error)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.host_group = host_group
        self.device = device
        self.flags = flags
        self.dim = lattice_dim(obj.bbox(), self.res)
        self.host_ms = {}   # wall time of the host-visible phases of the last run (diagnostics)
        import time
        t0 = time.perf_counter()
        if self.world > 1 and balanced:
            # the estimate is deterministic, every rank derives the same partition; it is computed
            # once per (scene, resolution, world) and cached
            self._part = cached_partition(ctx, obj, self.res, self.rank, self.world, group, device, host_group)
            if self._part.rounds < FEEDBACK_ROUNDS and self._part.my_last_us > 0:
                # measured-load refinement: every rank holds a device time of the previous run with these slabs
                # (report_device_time; all ranks report or none does, so they take this branch together)
                g, d = (host_group, "cpu") if host_group is not None else (group, device)
                times = [int(c[0]) for c in exchange_counts(self._part.my_last_us, 0, g, d)]
                refine_partition(self._part, times)
                self._part.my_last_us = 0
            self.slabs = self._part.slabs
        else:
            self._part = None
            self.slabs = partition(self.dim[2], self.world)
        self.host_ms["partition"] = (time.perf_counter() - t0) * 1e3
        if any(z1 == z0 for z0, z1 in self.slabs):
            raise ValueError("fewer cell layers than ranks")
        self.all_counts = None
        self.vertex_base = self.tet_base = 0

    def run(self):
        import time
        import torch
        lib = _lib.lib()
        t0 = time.perf_counter()

        def lap(name):
            nonlocal t0
            t1 = time.perf_counter()
            self.host_ms[name] = (t1 - t0) * 1e3
            t0 = t1

        z0, z1 = self.slabs[self.rank]
        slab = _lib.mc_slab(z0, z1)
        opts = _lib.mc_options(C.sizeof(_lib.mc_options), self.flags, 0, 0)
        out = C.c_void_p()
        _lib.check(lib.mc_tessellate_begin(self.ctx.handle, self.obj.backend.handle, self.obj.node, self.res, self.err,
                                           C.byref(slab), C.byref(opts), C.byref(out)))
        try:
            lap("begin")
            counts = (C.c_uint64 * 6)()
            _lib.check(lib.mc_mesh_counts(out, counts))
            top_first, top_count = C.c_uint64(), C.c_uint64()
            _lib.check(lib.mc_mesh_upper_layer_vertices(out, C.byref(top_first), C.byref(top_count)))
            extra = (top_first.value, top_count.value)
            if self.world > 1 and self.host_group is None:
                # device group: the stream is idle right after phase 1, the tiny all-gather is not
                # queued behind anything
                self.all_counts = exchange_counts(counts[0], counts[1], self.group, self.device, extra=extra)
            # coordinates + bisection do not need the global bases
            _lib.check(lib.mc_tessellate_emit_vertices_local(out))
            if self.world > 1 and self.host_group is not None:
                # host group: the device works on the vertices while the counts travel
                self.all_counts = exchange_counts(counts[0], counts[1], self.host_group, "cpu", extra=extra)
            elif self.world == 1:
                self.all_counts = [(counts[0], counts[1]) + extra]
            self.vertex_base, self.tet_base = bases_from_counts(self.all_counts, self.rank)
            lap("exchange_counts")
            _lib.check(lib.mc_tessellate_emit_vertices(out, self.vertex_base))
            lap("emit_vertices_launch")
            lower = None
            if self.world > 1:
                n = (self.dim[0] + 1) * (self.dim[1] + 1)
                send = recv = None
                if self.rank + 1 < self.world:
                    send = torch.empty(n, dtype=torch.int64, device=self.device)
                    _lib.check(lib.mc_mesh_copy_upper_plane_table(out, C.c_void_p(send.data_ptr())))
                send_xyz = recv_xyz = None
                if send is not None:
                    # coordinates of the top tile layer: the rank above evaluates the quality of the
                    # trimmed tets that touch the shared plane with them
                    send_xyz = torch.empty(3 * top_count.value, dtype=torch.float64, device=self.device)
                    if top_count.value:
                        _lib.check(lib.mc_mesh_copy_upper_layer_coords(out, C.c_void_p(send_xyz.data_ptr())))
                if self.rank > 0:
                    recv = torch.empty(n, dtype=torch.int64, device=self.device)
                    below = self.all_counts[self.rank - 1]
                    recv_xyz = torch.empty(3 * below[3], dtype=torch.float64, device=self.device)
                # The table / coordinate copies above were queued on the CONTEXT's stream and the tet
                # emission below reads the received table on it, while the NCCL point-to-point ops order
                # against torch's CURRENT stream: make the context's stream current for the exchange,
                # whatever stream the caller is on.
                with torch.cuda.stream(_ctx_stream(self.ctx, self.device)):
                    exchange_plane_tables(send, recv, self.rank, self.world, self.group, send_coords=send_xyz,
                                          recv_coords=recv_xyz)
                if recv is not None:
                    lower = C.c_void_p(recv.data_ptr())
                    vb_below, _ = bases_from_counts(self.all_counts, self.rank - 1)
                    _lib.check(lib.mc_tessellate_set_lower_coords(out, C.c_void_p(recv_xyz.data_ptr()) if below[3] else None,
                                                                  vb_below + below[2], below[3]))
                self._keep = (send, recv, send_xyz, recv_xyz)
            lap("exchange_plane_launch")
            rc = lib.mc_tessellate_emit_tets(out, self.tet_base, lower)
            lap("emit_tets")
            if rc not in (_lib.MC_OK, _lib.MC_ERR_DIVERGED):
                _lib.check(rc)
        except Exception:
            lib.mc_mesh_free(out)
            raise
        return out


def report_device_time(sm, device_ms):
    """Tell the partition cache how long this rank's device worked on its slab in the run `sm` just did
    (meshing + whatever the caller ran on the slab afterwards, e.g. Mesh::boundary): the next runs of
    the same scene rebalance the slabs with it (refine_partition)."""
    part = getattr(sm, "_part", None)
    if part is not None and part.slabs == sm.slabs:
        part.my_last_us = max(1, int(device_ms * 1000.0))


def tessellate_multi(ctxs, obj, resolution, relative_error, flags=0):
    """One process driving several devices (C ABI mc_tessellate_multi): ctxs[i] meshes slab i.
    Returns the list of raw mc_mesh handles (global vertex ids); free each with mc_mesh_free."""
    lib = _lib.lib()
    n = len(ctxs)
    handles = (C.c_void_p * n)(*[c.handle for c in ctxs])
    out = (C.c_void_p * n)()
    opts = _lib.mc_options(C.sizeof(_lib.mc_options), flags, 0, 0)
    rc = lib.mc_tessellate_multi(handles, n, obj.backend.handle, obj.node, float(resolution), float(relative_error),
                                 C.byref(opts), out)
    if rc not in (_lib.MC_OK, _lib.MC_ERR_DIVERGED):
        _lib.check(rc)
    return [C.c_void_p(h) for h in out]


def gather_slab_meshes(handles):
    """Host copy of slab meshes with global ids -> (coords (n, 3) f64, tets (m, 4) u64, combined stats)."""
    import numpy as np
    lib = _lib.lib()
    coords, tets, stats = [], [], []
    for h in handles:
        c = (C.c_uint64 * 6)()
        _lib.check(lib.mc_mesh_counts(h, c))
        cbuf = np.empty((c[0], 3), dtype=np.float64)
        tbuf = np.empty((c[1], 4), dtype=np.uint64)
        _lib.check(lib.mc_mesh_copy_to_host(h, cbuf.ctypes.data_as(C.c_void_p), tbuf.ctypes.data_as(C.c_void_p), 8))
        coords.append(cbuf)
        tets.append(tbuf)
        stats.append(mesh_stats(h))
    return np.concatenate(coords), np.concatenate(tets), combine_stats(stats)


def tessellate_slabs_one_device(ctx, obj, resolution, relative_error, n_slabs, flags=0, return_stats=False, id_offset=0):
    """Mesh the lattice as `n_slabs` z-slabs one after the other on ONE device and merge the
    results on the host (global vertex ids).  Exercises exactly the phases a multi-GPU run
    goes through (counts -> bases -> vertices -> plane table -> tets); also a way to mesh a
    lattice whose working set exceeds one GPU.  Returns (coords (n,3) f64, tets (m,4) u64).
    id_offset: the global vertex ids start there instead of at 0 (as if slabs with that many vertices lay
    below; ids >= 2^32 take the 64-bit index kernels)."""
    import numpy as np
    lib = _lib.lib()
    dim = lattice_dim(obj.bbox(), float(resolution))
    slabs = partition(dim[2], n_slabs)
    handles, counts = [], []
    try:
        for z0, z1 in slabs:
            slab = _lib.mc_slab(z0, z1)
            opts = _lib.mc_options(C.sizeof(_lib.mc_options), flags, 0, 0)
            out = C.c_void_p()
            _lib.check(lib.mc_tessellate_begin(ctx.handle, obj.backend.handle, obj.node, float(resolution),
                                               float(relative_error), C.byref(slab), C.byref(opts), C.byref(out)))
            handles.append(out)
            c = (C.c_uint64 * 6)()
            _lib.check(lib.mc_mesh_counts(out, c))
            counts.append((c[0], c[1]))
        lower = None
        coords, tets, stats = [], [], []
        for r, h in enumerate(handles):
            vb, tb = bases_from_counts(counts, r)
            vb += int(id_offset)
            _lib.check(lib.mc_tessellate_emit_vertices(h, vb))
            if r > 0:
                # coordinates of the slab below's top tile layer, straight from its device buffer
                first, cnt, xyz = C.c_uint64(), C.c_uint64(), C.c_void_p()
                _lib.check(lib.mc_mesh_upper_layer_vertices(handles[r - 1], C.byref(first), C.byref(cnt)))
                _lib.check(lib.mc_mesh_upper_layer_coords(handles[r - 1], C.byref(xyz)))
                vb_below, _ = bases_from_counts(counts, r - 1)
                _lib.check(lib.mc_tessellate_set_lower_coords(h, xyz, vb_below + int(id_offset) + first.value, cnt.value))
            rc = lib.mc_tessellate_emit_tets(h, tb, lower)
            if rc not in (_lib.MC_OK, _lib.MC_ERR_DIVERGED):
                _lib.check(rc)
            table, n = C.c_void_p(), C.c_uint64()
            _lib.check(lib.mc_mesh_upper_plane_table(h, C.byref(table), C.byref(n)))
            lower = table if n.value else None
            nv, nt = counts[r]
            cbuf = np.empty((nv, 3), dtype=np.float64)
            tbuf = np.empty((nt, 4), dtype=np.uint64)
            _lib.check(lib.mc_mesh_copy_to_host(h, cbuf.ctypes.data_as(C.c_void_p), tbuf.ctypes.data_as(C.c_void_p), 8))
            coords.append(cbuf)
            tets.append(tbuf)
            stats.append(mesh_stats(h))
    finally:
        for h in handles:
            lib.mc_mesh_free(h)
    if return_stats:
        return np.concatenate(coords), np.concatenate(tets), combine_stats(stats)
    return np.concatenate(coords), np.concatenate(tets)
