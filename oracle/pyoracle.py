"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp).

ctypes wrapper of liboracle.so: a scene backend with the same interface as
``moosecad_b200.luascad.Tape`` so that the same ``geom.*`` front-end (and the same scripts)
build the CPU restatement's object tree, plus the restated mesher.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None
_PD = C.POINTER(C.c_double)
_PU64 = C.POINTER(C.c_uint64)
_PI = C.POINTER(C.c_int)


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        V, D, I, U = C.c_void_p, C.c_double, C.c_int, C.c_uint64
        sig = {
            "orc_set_libm_mode": (None, [I]), "orc_get_libm_mode": (I, []),
            "orc_set_assumption": (I, [I, I]), "orc_num_assumptions": (I, []),
            "orc_exp": (D, [D]), "orc_ln": (D, [D]),
            "orc_scene_new": (V, []), "orc_scene_free": (None, [V]),
            "orc_plane": (I, [V, I, I, D]), "orc_plane_hesian": (I, [V, D, D, D, D]),
            "orc_plane_3_points": (I, [V, _PD, _PD, _PD]),
            "orc_sphere": (I, [V, D]), "orc_i_cylinder": (I, [V, D]), "orc_i_cone": (I, [V, D, D]),
            "orc_unit_sphere_sq": (I, [V]), "orc_set_bbox": (I, [V, I, _PD]),
            "orc_union": (I, [V, _PI, I, D]), "orc_intersection": (I, [V, _PI, I, D]),
            "orc_difference": (I, [V, _PI, I, D]), "orc_negation": (I, [V, I]),
            "orc_translate": (I, [V, I, D, D, D]), "orc_rotate": (I, [V, I, D, D, D]),
            "orc_scale": (I, [V, I, D, D, D]),
            "orc_set_parameters": (None, [V, I, D, D]), "orc_bbox": (None, [V, I, _PD]),
            "orc_approx_value": (D, [V, I, D, D, D, D]),
            "orc_eval_points": (None, [V, I, D, _PD, U, _PD]),
            "orc_eval_lattice": (None, [V, I, D, D, _PU64, _PD]),
            "orc_tessellate": (V, [V, I, D, D]), "orc_mesh_free": (None, [V]), "orc_mesh_ok": (I, [V]),
            "orc_mesh_num_vertices": (U, [V]), "orc_mesh_num_tets": (U, [V]),
            "orc_mesh_copy_coords": (None, [V, _PD]), "orc_mesh_copy_tets": (None, [V, _PU64]),
            "orc_mesh_stats": (None, [V, _PU64, _PD, _PD]),
            "orc_mesh_boundary": (U, [V]), "orc_mesh_copy_boundary": (None, [V, _PU64]),
            "orc_mesh_add_sideset": (I, [V, V, I]), "orc_mesh_sideset_len": (U, [V, I]),
            "orc_mesh_copy_sideset": (None, [V, I, _PU64]),
            "orc_rawmesh_new": (V, [_PD, U, _PU64, U]), "orc_rawmesh_compact": (None, [V]),
            "orc_rawmesh_volume": (D, [V, U]), "orc_rawmesh_aspect_ratio": (D, [V, U]),
            "orc_rawmesh_face_normal": (None, [V, U, I, _PD]),
            "orc_cut_edge_test": (C.c_int64, [V, I, D, D, _PD, _PD, _PD, _PD]),
            "orc_tile_lattice": (_PI, []), "orc_tile_tets": (_PI, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def set_libm_mode(mode):
    """0: platform libm (what the reference binary executes); 1: portable restatement (default)."""
    lib().orc_set_libm_mode(int(mode))


# Named assumptions about the un-vendored crates (enum Assumption in oracle.cpp).
# Value 0 is what the oracle and the device path implement, the others are the alternative readings.
ASSUMPTIONS = {
    "union_bbox_dilation": (0, {0: "r * 0.2f32 as f64", 1: "r * 0.2 (f64)", 2: "r", 3: "none"}),
    "intersection_bbox_dilation": (1, {0: "not dilated", 1: "dilated like the union"}),
    "euler_order": (2, {0: "Rz(yaw) Ry(pitch) Rx(roll)", 1: "Rx(roll) Ry(pitch) Rz(yaw)"}),
    "smooth_limit": (3, {0: "x < m + r enters the exponential sum", 1: "x < m + exact_range"}),
    "boolean_slack": (4, {0: "children get slack + r", 1: "slack unchanged"}),
}


def set_assumption(name, value):
    """Switch one named Appendix-A assumption of the oracle; returns the previous value.  Scenes must be
    built AFTER the switch (bounding boxes and matrices are computed at construction)."""
    idx, values = ASSUMPTIONS[name]
    if value not in values:
        raise ValueError(f"{name}: {value} not in {sorted(values)}")
    old = lib().orc_set_assumption(idx, int(value))
    if old < 0:
        raise RuntimeError("oracle rejected the assumption id")
    return old


def reset_assumptions():
    for name in ASSUMPTIONS:
        set_assumption(name, 0)


class OracleScene:
    """Scene backend (same methods as moosecad_b200.luascad.Tape) building the oracle's object tree."""

    def __init__(self):
        self._lib = lib()
        self.handle = self._lib.orc_scene_new()
        self._outputs = set()

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.orc_scene_free(h)

    @staticmethod
    def _ids(ids):
        return (C.c_int * len(ids))(*ids), len(ids)

    @staticmethod
    def _chk(v):
        if v < 0:
            raise RuntimeError(f"oracle scene error {v}")
        return v

    def plane(self, axis, negative, d): return self._chk(self._lib.orc_plane(self.handle, axis, int(negative), d))
    def plane_hesian(self, nx, ny, nz, p): return self._chk(self._lib.orc_plane_hesian(self.handle, nx, ny, nz, p))

    def plane_3_points(self, a, b, c):
        arr = lambda v: (C.c_double * 3)(*v)
        return self._chk(self._lib.orc_plane_3_points(self.handle, arr(a), arr(b), arr(c)))

    def sphere(self, r): return self._chk(self._lib.orc_sphere(self.handle, r))
    def i_cylinder(self, r): return self._chk(self._lib.orc_i_cylinder(self.handle, r))
    def i_cone(self, slope, offset): return self._chk(self._lib.orc_i_cone(self.handle, slope, offset))
    def unit_sphere_sq(self): return self._chk(self._lib.orc_unit_sphere_sq(self.handle))
    def set_bbox(self, node, bbox): self._lib.orc_set_bbox(self.handle, node, (C.c_double * 6)(*bbox))

    def union(self, ids, r):
        a, n = self._ids(ids)
        return self._chk(self._lib.orc_union(self.handle, a, n, r))

    def intersection(self, ids, r):
        a, n = self._ids(ids)
        return self._chk(self._lib.orc_intersection(self.handle, a, n, r))

    def difference(self, ids, r):
        a, n = self._ids(ids)
        return self._chk(self._lib.orc_difference(self.handle, a, n, r))

    def translate(self, node, x, y, z): return self._chk(self._lib.orc_translate(self.handle, node, x, y, z))
    def rotate(self, node, x, y, z): return self._chk(self._lib.orc_rotate(self.handle, node, x, y, z))
    def scale(self, node, x, y, z): return self._chk(self._lib.orc_scale(self.handle, node, x, y, z))
    def bend(self, node, w): raise NotImplementedError("bend is outside the path")
    def twist(self, node, h): raise NotImplementedError("twist is outside the path")

    def set_parameters_node(self, node, fade_range, r_multiplier):
        self._lib.orc_set_parameters(self.handle, node, fade_range, r_multiplier)

    def bbox(self, node):
        out = (C.c_double * 6)()
        self._lib.orc_bbox(self.handle, node, out)
        return list(out)

    # --- evaluation ---
    def approx_value(self, node, p, slack):
        return self._lib.orc_approx_value(self.handle, node, p[0], p[1], p[2], slack)

    def eval_points(self, node, slack, points):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pts), dtype=np.float64)
        self._lib.orc_eval_points(self.handle, node, slack, pts.ctypes.data_as(_PD), len(pts), out.ctypes.data_as(_PD))
        return out

    def eval_lattice(self, node, slack, res, dim):
        d = (C.c_uint64 * 3)(*dim)
        out = np.empty((dim[2] + 1, dim[1] + 1, dim[0] + 1), dtype=np.float64)
        self._lib.orc_eval_lattice(self.handle, node, slack, res, d, out.ctypes.data_as(_PD))
        return out

    def tessellate(self, node, resolution, relative_error):
        h = self._lib.orc_tessellate(self.handle, node, resolution, relative_error)
        return OracleMesh(h, self)

    def cut_edge_test(self, node, resolution, relative_error, v0, v1):
        pos = (C.c_double * 3)()
        phi = C.c_double()
        idx = self._lib.orc_cut_edge_test(self.handle, node, resolution, relative_error, (C.c_double * 3)(*v0),
                                          (C.c_double * 3)(*v1), pos, C.byref(phi))
        return idx, list(pos), phi.value


class OracleMesh:
    def __init__(self, handle, scene=None):
        self._lib = lib()
        self.handle = handle
        self.scene = scene

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.orc_mesh_free(h)

    @classmethod
    def from_arrays(cls, coords, tets):
        c = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
        t = np.ascontiguousarray(tets, dtype=np.uint64).reshape(-1, 4)
        return cls(lib().orc_rawmesh_new(c.ctypes.data_as(_PD), len(c), t.ctypes.data_as(_PU64), len(t)))

    def ok(self): return bool(self._lib.orc_mesh_ok(self.handle))

    def vertices(self):
        n = self._lib.orc_mesh_num_vertices(self.handle)
        out = np.empty((n, 3), dtype=np.float64)
        self._lib.orc_mesh_copy_coords(self.handle, out.ctypes.data_as(_PD))
        return out

    def elems(self):
        n = self._lib.orc_mesh_num_tets(self.handle)
        out = np.empty((n, 4), dtype=np.uint64)
        self._lib.orc_mesh_copy_tets(self.handle, out.ctypes.data_as(_PU64))
        return out

    def stats(self):
        u = (C.c_uint64 * 20)()
        ar = (C.c_double * 2)()
        t = (C.c_double * 8)()
        self._lib.orc_mesh_stats(self.handle, u, ar, t)
        u = list(u)
        return {"dim": u[0:3], "ntets_stage": u[3:6], "stats": u[6:13], "n_warped": u[13], "n_cut": u[14],
                "n_value_calls": u[15], "n_bisect_iters": u[16], "n_exterior_removed": u[17],
                "n_exterior_skipped_by_swap": u[18], "inverted": u[19], "aspect_ratio": list(ar),
                "times": dict(zip(["cut", "warp", "trim", "ext", "compact", "check", "boundary", "sidesets"], t))}

    def boundary(self):
        n = self._lib.orc_mesh_boundary(self.handle)
        out = np.empty((n, 2), dtype=np.uint64)
        self._lib.orc_mesh_copy_boundary(self.handle, out.ctypes.data_as(_PU64))
        return out

    def add_sideset(self, scene, node):
        i = self._lib.orc_mesh_add_sideset(self.handle, scene.handle, node)
        n = self._lib.orc_mesh_sideset_len(self.handle, i)
        out = np.empty((n, 2), dtype=np.uint64)
        self._lib.orc_mesh_copy_sideset(self.handle, i, out.ctypes.data_as(_PU64))
        return out

    def compact(self): self._lib.orc_rawmesh_compact(self.handle)
    def volume(self, e): return self._lib.orc_rawmesh_volume(self.handle, e)
    def aspect_ratio(self, e): return self._lib.orc_rawmesh_aspect_ratio(self.handle, e)

    def face_normal(self, e, s):
        out = (C.c_double * 3)()
        self._lib.orc_rawmesh_face_normal(self.handle, e, s, out)
        return np.array(list(out))
