"""Host-side mirror of ``tessellate3d`` + the parts of ``mesh`` on the accelerated path.

    IsosurfaceStuffing(function, resolution, relative_error).tessellate() -> Mesh
        tessellate3d/src/lib.rs:36-52,420-431 (called from src/main.rs:74-82)
    Mesh.vertices() / elems() / boundary() / add_side_set()
        mesh/src/lib.rs:58-65,138-140,183-185,341-357,198-202
    ObjectAdaptor(implicit, tessellation_resolution).value(points)
        src/main.rs:23-40

Everything runs on the GPU through the C ABI (``_lib``); there is no host fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from .luascad import LObject, Tape


class Context:
    """One ``mc_ctx``: a device, a stream and the reusable device workspace."""

    def __init__(self, device=0, stream=None):
        self._lib = _lib.lib()
        self.handle = self._lib.mc_ctx_new(int(device), C.c_void_p(stream) if stream else None)
        if not self.handle:
            raise _lib.McError(_lib.MC_ERR_CUDA, _lib.last_error())
        self.device = int(device)
        self.stream = int(stream) if stream else 0   # cudaStream_t the context launches on (0 = legacy default)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.mc_ctx_free(h)

    @property
    def launch_count(self):
        return int(self._lib.mc_ctx_launch_count(self.handle))

    @property
    def workspace_bytes(self):
        return int(self._lib.mc_ctx_workspace_bytes(self.handle))

    def trim(self):
        _lib.check(self._lib.mc_ctx_trim(self.handle))

    def probe_fp64_nofma(self):
        out = C.c_double()
        _lib.check(self._lib.mc_probe_fp64_nofma(self.handle, C.byref(out)))
        return out.value

    def tet_quality(self, coords, tets):
        """Mesh::volume / Mesh::aspect_ratio (mesh/src/lib.rs:292-316) for explicit arrays."""
        coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
        tets = np.ascontiguousarray(tets, dtype=np.uint64).reshape(-1, 4)
        vol = np.empty(len(tets), dtype=np.float64)
        ar = np.empty(len(tets), dtype=np.float64)
        _lib.check(self._lib.mc_tet_quality(
            self.handle, coords.ctypes.data_as(_lib._PD), len(coords), tets.ctypes.data_as(_lib._PU64), len(tets),
            vol.ctypes.data_as(_lib._PD), ar.ctypes.data_as(_lib._PD)))
        return vol, ar


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class ObjectAdaptor:
    """``ObjectAdaptor`` of src/main.rs:23-40: an implicit object + the slack passed to approx_value."""

    def __init__(self, implicit, tessellation_resolution, ctx=None):
        if not isinstance(implicit, LObject) or not isinstance(implicit.backend, Tape):
            raise TypeError("ObjectAdaptor needs an LObject recorded on a Tape")
        self.implicit = implicit
        self.tessellation_resolution = float(tessellation_resolution)
        self.ctx = ctx

    def bbox(self):
        return self.implicit.bbox()

    def value(self, points, slack=None):
        """approx_value(p, slack) for an (n, 3) array of points; evaluated on the device."""
        ctx = self.ctx or default_context()
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pts), dtype=np.float64)
        lib = _lib.lib()
        _lib.check(lib.mc_eval_points(ctx.handle, self.implicit.backend.handle, self.implicit.node,
                                      self.tessellation_resolution if slack is None else float(slack),
                                      pts.ctypes.data_as(_lib._PD), len(pts), out.ctypes.data_as(_lib._PD)))
        return out


class SideSet:
    """``mesh::SideSet`` (mesh/src/lib.rs:68-102): a named set of (elem, side) pairs.

    Large side-sets come back from the device as an (n, 2) uint64 array; the Python ``set`` of
    tuples the reference's ``HashSet`` corresponds to is only materialised when ``sides()`` is
    called.  ``array()`` gives the sorted (elem, side) rows without that cost."""

    def __init__(self, sides=(), name=None):
        self._name = name
        self._array = None
        self._sides = None
        if isinstance(sides, np.ndarray):
            self._array = np.ascontiguousarray(sides, dtype=np.uint64).reshape(-1, 2)
        else:
            self._sides = {(int(e), int(s)) for e, s in sides}

    def name(self):
        return self._name

    def set_name(self, name):
        self._name = str(name)

    def add_side(self, elem, side):
        s = self.sides()
        new = (int(elem), int(side)) not in s
        s.add((int(elem), int(side)))
        self._array = None
        return new

    def sides(self):
        if self._sides is None:
            self._sides = set(map(tuple, self._array.tolist()))
        return self._sides

    def array(self):
        if self._array is None:
            self._array = np.array(sorted(self._sides), dtype=np.uint64).reshape(-1, 2)
        return self._array

    def __len__(self):
        return len(self._array) if self._array is not None else len(self._sides)


class Mesh:
    """Result of a tessellation: ``Mesh<Tet4, f64>`` (mesh/src/lib.rs:58-65), device backed."""

    def __init__(self, ctx, handle, tape):
        self._lib = _lib.lib()
        self.ctx = ctx
        self.handle = handle
        self._tape = tape  # keeps the scene alive
        self._name = None
        self._side_sets = []
        self._coords = None
        self._tets = None

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.mc_mesh_free(h)

    def name(self):
        return self._name

    def set_name(self, name):
        self._name = str(name)

    def vertices(self):
        """(n, 3) float64 — ``Mesh::vertices`` (mesh/src/lib.rs:138-140)."""
        if self._coords is None:
            n = int(self._lib.mc_mesh_num_vertices(self.handle))
            p = self._lib.mc_mesh_coords(self.handle)
            if n and not p:
                raise _lib.McError(_lib.MC_ERR_CUDA, _lib.last_error())
            self._coords = np.ctypeslib.as_array(p, shape=(n, 3)).copy() if n else np.zeros((0, 3))
        return self._coords

    def elems(self):
        """(n, 4) uint64 node ids — ``Mesh::elems`` (mesh/src/lib.rs:183-185)."""
        if self._tets is None:
            n = int(self._lib.mc_mesh_num_tets(self.handle))
            p = self._lib.mc_mesh_tets(self.handle)
            if n and not p:
                raise _lib.McError(_lib.MC_ERR_CUDA, _lib.last_error())
            self._tets = np.ctypeslib.as_array(p, shape=(n, 4)).copy() if n else np.zeros((0, 4), dtype=np.uint64)
        return self._tets

    def num_vertices(self):
        return int(self._lib.mc_mesh_num_vertices(self.handle))

    def num_elems(self):
        return int(self._lib.mc_mesh_num_tets(self.handle))

    def boundary(self):
        """``Mesh::boundary`` (mesh/src/lib.rs:341-357): the sides that belong to one tet only."""
        n = C.c_uint64()
        _lib.check(self._lib.mc_mesh_boundary(self.handle, C.byref(n)))
        if n.value == 0:
            return SideSet()
        p = self._lib.mc_mesh_boundary_sides(self.handle)
        return SideSet(np.ctypeslib.as_array(p, shape=(n.value, 2)).copy())

    def to_boundary(self):
        """``Mesh::to_boundary`` (mesh/src/lib.rs:359-386): the surface as a triangle mesh,
        (vertices (n, 3) float64, triangles (m, 3) uint64), vertices compacted."""
        out = C.c_void_p()
        _lib.check(self._lib.mc_mesh_to_boundary(self.handle, C.byref(out)))
        try:
            nv, nt = int(self._lib.mc_surface_num_vertices(out)), int(self._lib.mc_surface_num_tris(out))
            v = np.ctypeslib.as_array(self._lib.mc_surface_coords(out), shape=(nv, 3)).copy() if nv else np.zeros((0, 3))
            t = np.ctypeslib.as_array(self._lib.mc_surface_tris(out), shape=(nt, 3)).copy() if nt else np.zeros((0, 3), dtype=np.uint64)
        finally:
            self._lib.mc_surface_free(out)
        return v, t

    def add_boundary_side_set(self, name, boundary_obj):
        """One iteration of the side-set loop of src/main.rs:87-106 for a ``:boundary()`` object."""
        idx = _lib.check(self._lib.mc_mesh_add_sideset(self.handle, boundary_obj.backend.handle, boundary_obj.node))
        n = int(self._lib.mc_mesh_sideset_len(self.handle, idx))
        sides = np.zeros((0, 2), dtype=np.uint64)
        if n:
            sides = np.ctypeslib.as_array(self._lib.mc_mesh_sideset(self.handle, idx), shape=(n, 2)).copy()
        ss = SideSet(sides, name)
        self._side_sets.append(ss)
        return ss

    def add_side_set(self, side_set):
        self._side_sets.append(side_set)
        return len(self._side_sets) - 1

    def side_sets(self):
        return self._side_sets

    def stats(self):
        """The numbers the reference prints (tessellate3d/src/lib.rs:247,355,366,408-417,421,424)."""
        dim = (C.c_uint64 * 3)()
        stage = (C.c_uint64 * 3)()
        st = (C.c_uint64 * 7)()
        ar = (C.c_double * 2)()
        inv = C.c_uint64()
        _lib.check(self._lib.mc_mesh_stats(self.handle, dim, stage, st, ar, C.byref(inv)))
        counts = (C.c_uint64 * 6)()
        _lib.check(self._lib.mc_mesh_counts(self.handle, counts))
        ms = (C.c_double * 9)()
        _lib.check(self._lib.mc_mesh_stage_ms(self.handle, ms))
        nt, pw, fw = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _lib.check(self._lib.mc_mesh_prune_stats(self.handle, C.byref(nt), C.byref(pw), C.byref(fw)))
        ts = (C.c_uint64 * 6)()
        _lib.check(self._lib.mc_mesh_tile_stats(self.handle, ts))
        return {"tiles": {"total": ts[0], "sdf_skipped": ts[1], "active_point_tiles": ts[2], "active_cell_tiles": ts[3], "sdf_flops_executed": ts[4], "sub_boxes_proven": ts[5]},
                "prune": {"tiles": nt.value, "avg_words": (pw.value / nt.value) if nt.value else None,
                          "full_words": fw.value},
                "dim": list(dim), "ntets_stage": list(stage), "stats": list(st), "aspect_ratio": list(ar),
                "inverted": inv.value, "n_vertices": counts[0], "n_tets": counts[1],
                "points_evaluated": counts[2], "points_owned": counts[3], "n_cut": counts[4],
                "n_warped": counts[5], "stage_ms": list(ms)}

    def exodus_arrays(self):
        """The arrays ``exodus::write`` builds from the mesh (exodus/src/lib.rs:55-63,84-96,118-132), laid
        out on the device: {"coord": (3, num_nodes) f64, "connect1": (num_elem, 4) i64 1-based,
        "side_sets": [(elem_ss i32, side_ss i32) 1-based per side-set added so far]}."""
        nv, nt = self.num_vertices(), self.num_elems()
        coord = np.empty((3, nv), dtype=np.float64)
        connect = np.empty((nt, 4), dtype=np.int64)
        _lib.check(self._lib.mc_mesh_exodus_coord(self.handle, coord.ctypes.data_as(C.c_void_p), nv))
        _lib.check(self._lib.mc_mesh_exodus_connect(self.handle, connect.ctypes.data_as(C.c_void_p)))
        sets = []
        for i in range(len(self._side_sets)):
            n = int(self._lib.mc_mesh_sideset_len(self.handle, i))
            e, s = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
            _lib.check(self._lib.mc_mesh_exodus_sideset(self.handle, i, e.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p)))
            sets.append((e, s))
        return {"coord": coord, "connect1": connect, "side_sets": sets}

    def checksum(self):
        """Order- and numbering-independent checksums of the emitted mesh, computed on the device
        (C ABI mc_mesh_checksum): {"vertices": u64, "tets": u64, "tets_skipped": n}."""
        return mesh_checksum(self.handle)

    def phi(self):
        """The lattice SDF field, shape (nz+1, ny+1, nx+1)."""
        d = self.stats()["dim"]
        shape = (d[2] + 1, d[1] + 1, d[0] + 1)
        out = np.empty(shape, dtype=np.float64)
        _lib.check(self._lib.mc_mesh_copy_phi(self.handle, out.ctypes.data_as(_lib._PD), out.size))
        return out


def mesh_checksum(handle):
    """mc_mesh_checksum for a raw mc_mesh handle."""
    out = (C.c_uint64 * 4)()
    _lib.check(_lib.lib().mc_mesh_checksum(handle, out))
    return {"vertices": int(out[0]), "tets": int(out[1]), "tets_skipped": int(out[2])}


class IsosurfaceStuffing:
    """``tessellate3d::IsosurfaceStuffing`` (tessellate3d/src/lib.rs:24-52,420-431)."""

    def __init__(self, function, resolution, relative_error, ctx=None, flags=0, max_bisect_iters=0,
                 allow_diverged=False):
        if isinstance(function, LObject):
            function = ObjectAdaptor(function, resolution, ctx)
        if not isinstance(function, ObjectAdaptor):
            raise TypeError("function must be an ObjectAdaptor or an LObject")
        if function.tessellation_resolution != float(resolution):
            # src/main.rs:74-81 always passes the same number for both
            raise ValueError("the accelerated path evaluates approx_value with slack == resolution")
        self.function = function
        self.resolution = float(resolution)
        self.relative_error = float(relative_error)
        self.ctx = ctx or function.ctx or default_context()
        self.flags = int(flags)
        self.max_bisect_iters = int(max_bisect_iters)
        # the reference's cut_edge loops forever when the SDF jumps across zero by more than the
        # error threshold; here that is MC_ERR_DIVERGED (the mesh is still returned on request)
        self.allow_diverged = bool(allow_diverged)

    def tessellate(self):
        lib = _lib.lib()
        opts = _lib.mc_options(C.sizeof(_lib.mc_options), self.flags, self.max_bisect_iters, 0)
        out = C.c_void_p()
        obj = self.function.implicit
        rc = lib.mc_tessellate(self.ctx.handle, obj.backend.handle, obj.node, self.resolution, self.relative_error,
                               C.byref(opts), C.byref(out))
        if rc != _lib.MC_OK and not (rc == _lib.MC_ERR_DIVERGED and self.allow_diverged and out.value):
            err = _lib.McError(rc, _lib.last_error())
            if out.value:
                lib.mc_mesh_free(out)
            raise err
        return Mesh(self.ctx, out, obj.backend)


class Tile:
    """``tessellate3d::Tile`` (tessellate3d/src/lib.rs:434-454)."""
    LATTICE = [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]]
    TETS = [[0, 1, 5, 2], [0, 2, 5, 7], [0, 7, 5, 4], [2, 6, 5, 7], [0, 2, 7, 3]]
