"""ctypes binding of libmoosecad_b200.so (include/moosecad_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C moosecad_b200/csrc``.
There is no fallback of any kind: if the shared library is missing, or no CUDA device is
usable when a compute entry point is called, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmoosecad_b200.so")

MC_OK = 0
MC_ERR_ARG, MC_ERR_BBOX, MC_ERR_SINGULAR, MC_ERR_UNSUPPORTED = -1, -2, -3, -4
MC_ERR_CUDA, MC_ERR_LIMIT, MC_ERR_DIVERGED, MC_ERR_STATE = -5, -6, -7, -8
MC_OPT_KEEP_PHI, MC_OPT_SKIP_QUALITY, MC_OPT_NO_TILE_PRUNING, MC_OPT_FORCE_TILE_PRUNING = 1, 2, 4, 8
MC_OPT_COUNT_FLOPS = 16
MC_OPT_PER_TILE_QUALITY = 32
MC_OPT_RECLASSIFY_EMIT = 64
MC_TILE_CLASS_STRIDE = 7


class McError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"moosecad_b200 error {code}: {msg}")
        self.code = code


class mc_slab(C.Structure):
    _fields_ = [("z_begin", C.c_uint64), ("z_end", C.c_uint64)]


class mc_options(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32),
                ("max_bisect_iters", C.c_uint32), ("reserved", C.c_uint32)]


_P = C.c_void_p
_D = C.c_double
_I32 = C.c_int32
_U64 = C.c_uint64
_PD = C.POINTER(C.c_double)
_PU64 = C.POINTER(C.c_uint64)
_PI32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol declared in include/moosecad_b200.h
SIGNATURES = {
    "mc_last_error": (C.c_char_p, []),
    "mc_abi_version": (C.c_int, []),
    "mc_tape_new": (_P, []),
    "mc_tape_free": (None, [_P]),
    "mc_tape_num_nodes": (_I32, [_P]),
    "mc_tape_plane": (_I32, [_P, C.c_int, C.c_int, _D]),
    "mc_tape_plane_hesian": (_I32, [_P, _D, _D, _D, _D]),
    "mc_tape_plane_3_points": (_I32, [_P, _PD, _PD, _PD]),
    "mc_tape_sphere": (_I32, [_P, _D]),
    "mc_tape_i_cylinder": (_I32, [_P, _D]),
    "mc_tape_i_cone": (_I32, [_P, _D, _D]),
    "mc_tape_set_bbox": (C.c_int, [_P, _I32, _PD]),
    "mc_tape_unit_sphere_sq": (_I32, [_P]),
    "mc_tape_union": (_I32, [_P, _PI32, C.c_int, _D]),
    "mc_tape_intersection": (_I32, [_P, _PI32, C.c_int, _D]),
    "mc_tape_difference": (_I32, [_P, _PI32, C.c_int, _D]),
    "mc_tape_negation": (_I32, [_P, _I32]),
    "mc_tape_translate": (_I32, [_P, _I32, _D, _D, _D]),
    "mc_tape_rotate": (_I32, [_P, _I32, _D, _D, _D]),
    "mc_tape_scale": (_I32, [_P, _I32, _D, _D, _D]),
    "mc_tape_bend": (_I32, [_P, _I32, _D]),
    "mc_tape_twist": (_I32, [_P, _I32, _D]),
    "mc_tape_set_parameters": (C.c_int, [_P, _D, _D]),
    "mc_tape_bbox": (C.c_int, [_P, _I32, _PD]),
    "mc_tape_program_info": (C.c_int, [_P, _I32, _D, _PU64, _PI32, _PI32, _PU64]),
    "mc_tape_program_hash": (C.c_int, [_P, _I32, _D, _PU64]),
    "mc_tape_program_words": (C.c_int64, [_P, _I32, _D, _PU64, C.c_uint64]),
    "mc_ctx_new": (_P, [C.c_int, _P]),
    "mc_ctx_free": (None, [_P]),
    "mc_ctx_device": (C.c_int, [_P]),
    "mc_ctx_launch_count": (_U64, [_P]),
    "mc_ctx_workspace_bytes": (_U64, [_P]),
    "mc_ctx_trim": (C.c_int, [_P]),
    "mc_eval_points": (C.c_int, [_P, _P, _I32, _D, _PD, _U64, _PD]),
    "mc_eval_points_device": (C.c_int, [_P, _P, _I32, _D, _P, _U64, _P]),
    "mc_estimate_layer_costs": (C.c_int, [_P, _P, _I32, _D, _PD, _U64]),
    "mc_tile_layer_classes": (C.c_int, [_P, _P, _I32, _D, _U64, _U64, C.POINTER(_U64), _U64]),
    "mc_slab_class_weights": (None, [_PD]),
    "mc_tessellate_begin": (C.c_int, [_P, _P, _I32, _D, _D, C.POINTER(mc_slab), C.POINTER(mc_options),
                                      C.POINTER(_P)]),
    "mc_tessellate_emit_vertices": (C.c_int, [_P, _U64]),
    "mc_tessellate_emit_vertices_local": (C.c_int, [_P]),
    "mc_tessellate_set_lower_coords": (C.c_int, [_P, _P, _U64, _U64]),
    "mc_mesh_upper_layer_vertices": (C.c_int, [_P, C.POINTER(_U64), C.POINTER(_U64)]),
    "mc_mesh_upper_layer_coords": (C.c_int, [_P, C.POINTER(_P)]),
    "mc_mesh_copy_upper_layer_coords": (C.c_int, [_P, _P]),
    "mc_tessellate_emit_tets": (C.c_int, [_P, _U64, _P]),
    "mc_tessellate": (C.c_int, [_P, _P, _I32, _D, _D, C.POINTER(mc_options), C.POINTER(_P)]),
    "mc_tessellate_multi": (C.c_int, [C.POINTER(_P), C.c_int, _P, _I32, _D, _D, C.POINTER(mc_options), C.POINTER(_P)]),
    "mc_mesh_free": (None, [_P]),
    "mc_mesh_counts": (C.c_int, [_P, _PU64]),
    "mc_mesh_prune_stats": (C.c_int, [_P, _PU64, _PU64, _PU64]),
    "mc_mesh_tile_stats": (C.c_int, [_P, _PU64]),
    "mc_mesh_upper_plane_table": (C.c_int, [_P, C.POINTER(_P), _PU64]),
    "mc_mesh_copy_upper_plane_table": (C.c_int, [_P, _P]),
    "mc_mesh_stats": (C.c_int, [_P, _PU64, _PU64, _PU64, _PD, _PU64]),
    "mc_mesh_stage_ms": (C.c_int, [_P, _PD]),
    "mc_mesh_num_vertices": (_U64, [_P]),
    "mc_mesh_num_tets": (_U64, [_P]),
    "mc_mesh_coords": (_PD, [_P]),
    "mc_mesh_tets": (_PU64, [_P]),
    "mc_mesh_device_views": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int)]),
    "mc_mesh_copy_to_host": (C.c_int, [_P, _P, _P, C.c_int]),
    "mc_mesh_copy_phi": (C.c_int, [_P, _PD, _U64]),
    "mc_mesh_checksum": (C.c_int, [_P, _PU64]),
    "mc_mesh_exodus_coord": (C.c_int, [_P, _P, _U64]),
    "mc_mesh_exodus_connect": (C.c_int, [_P, _P]),
    "mc_mesh_exodus_sideset": (C.c_int, [_P, C.c_int, _P, _P]),
    "mc_mesh_to_boundary": (C.c_int, [_P, C.POINTER(_P)]),
    "mc_surface_free": (None, [_P]),
    "mc_surface_num_vertices": (_U64, [_P]),
    "mc_surface_num_tris": (_U64, [_P]),
    "mc_surface_coords": (_PD, [_P]),
    "mc_surface_tris": (_PU64, [_P]),
    "mc_mesh_boundary": (C.c_int, [_P, _PU64]),
    "mc_mesh_boundary_sides": (_PU64, [_P]),
    "mc_mesh_boundary_device": (C.c_int, [_P, C.POINTER(_P), _PU64]),
    "mc_mesh_sideset_device": (C.c_int, [_P, C.c_int, C.POINTER(_P), _PU64]),
    "mc_mesh_add_sideset": (C.c_int, [_P, _P, _I32]),
    "mc_mesh_sideset_len": (_U64, [_P, C.c_int]),
    "mc_mesh_sideset": (_PU64, [_P, C.c_int]),
    "mc_tet_quality": (C.c_int, [_P, _PD, _U64, _PU64, _U64, _PD, _PD]),
    "mc_probe_fp64_nofma": (C.c_int, [_P, _PD]),
    "mc_shape_tables": (C.c_int64, [_P, C.c_uint64]),
}

_lib = None


def lib():
    """The loaded shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(moosecad_b200 has no CPU / pure-Python fallback)")
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    return lib().mc_last_error().decode("utf-8", "replace")


def check(rc):
    """Raise McError for a negative status; return rc otherwise (node ids are >= 0)."""
    if rc < 0:
        raise McError(rc, last_error())
    return rc
