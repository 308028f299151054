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


def bases_from_counts(all_counts, rank):
    """Exclusive scan over ranks of (n_vertices, n_tets)."""
    vb = sum(int(c[0]) for c in all_counts[:rank])
    tb = sum(int(c[1]) for c in all_counts[:rank])
    return vb, tb


def exchange_counts(n_vertices, n_tets, group=None, device="cpu"):
    """all-gather of two int64 per rank; returns a list of (n_vertices, n_tets)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.tensor([int(n_vertices), int(n_tets)], dtype=torch.int64, device=device)
    out = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [(int(t[0]), int(t[1])) for t in out]


def exchange_plane_tables(send_table, recv_table, rank, world, group=None, active=None):
    """Rank r sends its upper-plane id table to r + 1 and receives r - 1's.  `send_table` /
    `recv_table` are int64 tensors (or None at the ends).  `active[r]` tells whether rank r
    owns any cell layer; empty ranks are skipped by relaying nothing."""
    import torch.distributed as dist
    ops = []
    if send_table is not None and rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, send_table, rank + 1, group))
    if recv_table is not None and rank > 0:
        ops.append(dist.P2POp(dist.irecv, recv_table, rank - 1, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def tile_layer_classes(ctx, obj, resolution):
    """Tile counts per tile layer and cost class (C ABI mc_tile_layer_classes): (n_tile_layers, 5)."""
    import numpy as np
    dim = lattice_dim(obj.bbox(), float(resolution))
    n = (dim[2] + 1 + 15) // 16
    counts = np.zeros((n, 5), dtype=np.uint64)
    _lib.check(_lib.lib().mc_tile_layer_classes(ctx.handle, obj.backend.handle, obj.node, float(resolution),
                                                counts.ctypes.data_as(C.POINTER(C.c_uint64)), n))
    return counts


class SlabMesher:
    """Meshes this rank's slab of the global lattice.  Usage (every rank):

        sm = SlabMesher(ctx, obj, res, err, rank, world, group, device)
        mesh_handle = sm.run()        # device-resident slab mesh with GLOBAL vertex ids
    """

    def __init__(self, ctx, obj, resolution, relative_error, rank, world, group=None, device=None, flags=0,
                 balanced=True):
        self.ctx, self.obj = ctx, obj
        self.res, self.err = float(resolution), float(relative_error)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = device
        self.flags = flags
        self.dim = lattice_dim(obj.bbox(), self.res)
        if self.world > 1 and balanced:
            # the estimate is deterministic, every rank derives the same partition
            self.slabs = partition_balanced(layer_costs(ctx, obj, self.res).tolist(), self.world)
        else:
            self.slabs = partition(self.dim[2], self.world)
        if any(z1 == z0 for z0, z1 in self.slabs):
            raise ValueError("fewer cell layers than ranks")
        self.all_counts = None
        self.vertex_base = self.tet_base = 0

    def run(self):
        import torch
        lib = _lib.lib()
        z0, z1 = self.slabs[self.rank]
        slab = _lib.mc_slab(z0, z1)
        opts = _lib.mc_options(C.sizeof(_lib.mc_options), self.flags, 0, 0)
        out = C.c_void_p()
        _lib.check(lib.mc_tessellate_begin(self.ctx.handle, self.obj.backend.handle, self.obj.node, self.res, self.err,
                                           C.byref(slab), C.byref(opts), C.byref(out)))
        try:
            counts = (C.c_uint64 * 6)()
            _lib.check(lib.mc_mesh_counts(out, counts))
            if self.world > 1:
                self.all_counts = exchange_counts(counts[0], counts[1], self.group, self.device)
            else:
                self.all_counts = [(counts[0], counts[1])]
            self.vertex_base, self.tet_base = bases_from_counts(self.all_counts, self.rank)
            _lib.check(lib.mc_tessellate_emit_vertices(out, self.vertex_base))
            lower = None
            if self.world > 1:
                n = (self.dim[0] + 1) * (self.dim[1] + 1)
                send = recv = None
                if self.rank + 1 < self.world:
                    send = torch.empty(n, dtype=torch.int64, device=self.device)
                    _lib.check(lib.mc_mesh_copy_upper_plane_table(out, C.c_void_p(send.data_ptr())))
                if self.rank > 0:
                    recv = torch.empty(n, dtype=torch.int64, device=self.device)
                exchange_plane_tables(send, recv, self.rank, self.world, self.group)
                if recv is not None:
                    lower = C.c_void_p(recv.data_ptr())
                self._keep = (send, recv)
            rc = lib.mc_tessellate_emit_tets(out, self.tet_base, lower)
            if rc not in (_lib.MC_OK, _lib.MC_ERR_DIVERGED):
                _lib.check(rc)
        except Exception:
            lib.mc_mesh_free(out)
            raise
        return out


def tessellate_slabs_one_device(ctx, obj, resolution, relative_error, n_slabs, flags=0):
    """Mesh the lattice as `n_slabs` z-slabs one after the other on ONE device and merge the
    results on the host (global vertex ids).  Exercises exactly the phases a multi-GPU run
    goes through (counts -> bases -> vertices -> plane table -> tets); also a way to mesh a
    lattice whose working set exceeds one GPU.  Returns (coords (n,3) f64, tets (m,4) u64)."""
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
        coords, tets = [], []
        for r, h in enumerate(handles):
            vb, tb = bases_from_counts(counts, r)
            _lib.check(lib.mc_tessellate_emit_vertices(h, vb))
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
    finally:
        for h in handles:
            lib.mc_mesh_free(h)
    return np.concatenate(coords), np.concatenate(tets)
