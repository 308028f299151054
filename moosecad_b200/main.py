"""The flow of the reference's ``src/main.rs`` around the accelerated path.

``run(script, ...)`` does what ``moosecad -i scene.lua ...`` does between reading the file and
calling ``exodus::write`` (src/main.rs:45-106): evaluate the script, apply
``set_parameters`` to every output, tessellate the FIRST volume, take ``mesh.boundary()`` and
build one side-set per ``:boundary()`` output.  Flag defaults are the reference's
(src/main.rs:45-48).  Writing the Exodus file stays with the reference's ``exodus`` crate.
"""
from . import luascad
from .tessellate3d import IsosurfaceStuffing, ObjectAdaptor

DEFAULT_FADE_RANGE = 0.1
DEFAULT_R_MULTIPLIER = 1.0
DEFAULT_TESSELLATION_RESOLUTION = 0.12
DEFAULT_TESSELLATION_ERROR = 0.2


def run(script, fade_range=None, r_multiplier=None, tessellation_resolution=None, tessellation_error=None,
        ctx=None, verbose=False, flags=0):
    fade_range = DEFAULT_FADE_RANGE if fade_range is None else fade_range
    r_multiplier = DEFAULT_R_MULTIPLIER if r_multiplier is None else r_multiplier
    res = DEFAULT_TESSELLATION_RESOLUTION if tessellation_resolution is None else tessellation_resolution
    err = DEFAULT_TESSELLATION_ERROR if tessellation_error is None else tessellation_error

    output = luascad.eval(script)
    tapes = {id(o.object_mut().backend): o.object_mut().backend for o in output.values()}
    for tape in tapes.values():  # src/main.rs:53-58
        tape.set_parameters(fade_range, r_multiplier)
    # the reference iterates a HashMap (arbitrary order); sorted order keeps runs reproducible
    volumes = [(n, o.volume()) for n, o in sorted(output.items()) if o.volume() is not None]
    boundaries = [(n, o.boundary()) for n, o in sorted(output.items()) if o.boundary() is not None]
    if not volumes:
        if verbose:
            print("no volumes")
        return None
    name, vol = volumes[0]
    if verbose:
        print(f"tessellating {name}")
    mesh = IsosurfaceStuffing(ObjectAdaptor(vol, res, ctx), res, err, ctx=ctx, flags=flags).tessellate()
    mesh.set_name(name)
    if verbose:
        st = mesh.stats()
        print(f"dim {st['dim']}")
        print(f"num tets {st['ntets_stage'][0]}")
        print(f"num tets {st['ntets_stage'][1]}")
        print(f"stats {st['stats']}")
        print(f"num tets {st['ntets_stage'][2]}")
        print("checking for inverted tets")
        print(f"aspect_ratio: min = {st['aspect_ratio'][0]} max = {st['aspect_ratio'][1]}")
        print("tessellation complete")
    for bname, bobj in boundaries:  # src/main.rs:86-106
        if verbose:
            print(f"adding boundary {bname}")
        mesh.add_boundary_side_set(bname, bobj)
    return mesh
