"""moosecad_b200 — B200-native drop-in for moosecad's meshing hot path.

Evaluating the luascad CSG signed-distance tree over tessellate3d's 5-tet cubic lattice and
running isosurface stuffing on it, as hand-written sm_100a CUDA behind a C ABI
(include/moosecad_b200.h).  The modules mirror the reference crates on that path:

    luascad        geom.* constructors / methods, eval()        (luascad/src/lib.rs)
    tessellate3d   IsosurfaceStuffing, ObjectAdaptor, Mesh      (tessellate3d/src/lib.rs, mesh/src/lib.rs)
    main           the flow of src/main.rs (first volume + boundary side-sets)
    scenes         the synthetic CSG scenes BASELINE.json names
    slabs          z-slab sharding over torch.distributed ranks

Importing the package does not need a GPU; every compute call does (no CPU fallback).
"""
from . import _lib  # noqa: F401
from .luascad import Geom, LObject, Output, Tape, LuaError  # noqa: F401
from .luascad import eval as eval_script  # noqa: F401
from .tessellate3d import Context, IsosurfaceStuffing, Mesh, ObjectAdaptor, SideSet, Tile  # noqa: F401

__all__ = ["Geom", "LObject", "Output", "Tape", "LuaError", "eval_script", "Context", "IsosurfaceStuffing",
           "Mesh", "ObjectAdaptor", "SideSet", "Tile"]
