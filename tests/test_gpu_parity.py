"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs.
Bit-exact for SDF values (oracle libm_mode 1 = the portable exp/ln the device implements),
vertex coordinates, tet connectivity and side-sets after canonical sorting."""
import math

import numpy as np
import pytest

import moosecad_b200 as mc
from moosecad_b200 import _lib, luascad, main, scenes

import canon

pytestmark = pytest.mark.gpu


def _oracle_run(oracle, script, res, err=0.2, fade=0.1, rmul=1.0):
    out = luascad.eval(script, backend=oracle.OracleScene())
    for o in out.values():
        o.obj.backend.set_parameters_node(o.obj.node, fade, rmul)
    vols = [(n, o.volume()) for n, o in sorted(out.items()) if o.volume() is not None]
    bnds = [(n, o.boundary()) for n, o in sorted(out.items()) if o.boundary() is not None]
    v = vols[0][1]
    om = v.backend.tessellate(v.node, res, err)
    assert om.ok()
    sets = {n: om.add_sideset(b.backend, b.node) for n, b in bnds}
    return om, sets, v


def _compare(oracle, ctx, script, res, err=0.2, bitwise=True, flags=0):
    mesh = main.run(script, tessellation_resolution=res, tessellation_error=err, ctx=ctx, flags=flags)
    om, osets, _ = _oracle_run(oracle, script, res, err)
    gsets = {ss.name(): sorted(ss.sides()) for ss in mesh.side_sets()}
    a = canon.canonical(mesh.vertices(), mesh.elems(), gsets)
    b = canon.canonical(om.vertices(), om.elems(), osets)
    canon.assert_same_mesh(a, b, bitwise=bitwise)
    st, ost = mesh.stats(), om.stats()
    assert st["dim"] == ost["dim"]
    assert st["ntets_stage"] == ost["ntets_stage"]
    assert st["stats"] == ost["stats"]
    assert st["n_warped"] == ost["n_warped"]
    assert st["n_cut"] == ost["n_cut"]
    assert st["inverted"] == ost["inverted"]
    assert ost["n_exterior_skipped_by_swap"] == 0  # precondition of the set-semantics restatement
    np.testing.assert_allclose(st["aspect_ratio"], ost["aspect_ratio"], rtol=1e-12)
    # Mesh::boundary as a set of node triples
    gb = canon.canonical(mesh.vertices(), mesh.elems(), {"surface": sorted(mesh.boundary().sides())})[2]["surface"]
    ob = canon.canonical(om.vertices(), om.elems(), {"surface": om.boundary()})[2]["surface"]
    assert np.array_equal(gb, ob)
    return mesh, om


def test_luascad_primitives_asserts(ctx):
    # luascad/src/lib.rs:375-377, evaluated on the device
    res = luascad.eval(scenes.CUBE_LUA)
    top = res["top"].boundary()
    v = mc.ObjectAdaptor(top, 0.0, ctx).value([[0.0, 0.0, 0.5], [0.0, 0.0, 1.0]])
    assert v[0] == 0.0 and v[1] == 0.5


SCRIPTS = {
    "cube": scenes.CUBE_LUA,
    "sphere": scenes.SPHERE_LUA,
    "xplicit": scenes.XPLICIT_LUA,
    "cyl_cone": "return { c = geom.cylinder(2.0, 0.5, 0.25, 0.1):rotate(0.3, 0.2, 0.1):translate(1, 2, 3):scale(2, 3, 4):volume() }",
    "cyl": "return { c = geom.cylinder(1.0, 0.5, 0.5, 0.05):translate(0.1, 0.2, 0.3):volume() }",
    "mixed": "a = geom.sphere(0.3):translate(0.2, 0, 0)\nb = geom.cube(0.5, 0.4, 0.3, 0.05):rotate(1, 2, 3)\n"
             "c = geom.plane_hesian(1, 2, 3, 0.1)\nd = geom.plane_3_points(0,0,0.2, 1,0,0.25, 0,1,0.3)\n"
             "return { u = a:union(b, 0.1):intersection(c, 0.05):difference(d, 0.0):volume() }",
}


@pytest.mark.parametrize("name", sorted(SCRIPTS))
def test_sdf_values_bit_exact(name, oracle, ctx):
    a = luascad.eval(SCRIPTS[name])
    b = luascad.eval(SCRIPTS[name], backend=oracle.OracleScene())
    rng = np.random.default_rng(1)
    for k in a:
        bb = np.array(a[k].obj.bbox())
        lo = np.where(np.isfinite(bb[:3]), bb[:3], -2.0) - 0.3
        hi = np.where(np.isfinite(bb[3:]), bb[3:], 2.0) + 0.3
        pts = rng.uniform(lo, hi, size=(20000, 3))
        pts[:100] = 0.0
        pts[100:200, 0] = bb[0] if np.isfinite(bb[0]) else 0.0
        for slack in (0.0, 0.05, 1.0):
            got = mc.ObjectAdaptor(a[k].obj, slack, ctx).value(pts)
            want = b[k].obj.backend.eval_points(b[k].obj.node, slack, pts)
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)) or np.array_equal(got, want), \
                (name, k, slack, np.abs(got - want).max())


def test_sdf_values_vs_platform_libm(oracle, ctx):
    # the reference calls the platform libm; the device's portable exp/ln stays within the
    # north-star tolerance (1e-12 relative) of it
    a = luascad.eval(scenes.XPLICIT_LUA)["xplicit"].obj
    b = luascad.eval(scenes.XPLICIT_LUA, backend=oracle.OracleScene())["xplicit"].obj
    rng = np.random.default_rng(2)
    pts = rng.uniform(-8, 8, size=(50000, 3))
    got = mc.ObjectAdaptor(a, 0.05, ctx).value(pts)
    oracle.set_libm_mode(0)
    try:
        want = b.backend.eval_points(b.node, 0.05, pts)
    finally:
        oracle.set_libm_mode(1)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)


def test_lattice_field_bit_exact(oracle, ctx):
    res = 15.0 / 24
    a = luascad.eval(scenes.XPLICIT_LUA)["xplicit"].obj
    mesh = mc.IsosurfaceStuffing(a, res, 0.2, ctx=ctx, flags=_lib.MC_OPT_KEEP_PHI).tessellate()
    b = luascad.eval(scenes.XPLICIT_LUA, backend=oracle.OracleScene())["xplicit"].obj
    want = b.backend.eval_lattice(b.node, res, res, mesh.stats()["dim"])
    got = mesh.phi()
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_config1_cube_lua(oracle, ctx):
    mesh, om = _compare(oracle, ctx, scenes.CUBE_LUA, 1.0)
    assert mesh.num_vertices() == 8 and mesh.num_elems() == 5
    assert len(mesh.boundary().sides()) == 12
    assert {ss.name(): len(ss.sides()) for ss in mesh.side_sets()} == {"top": 2, "bottom": 2}


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 33, 64])
def test_sphere_lua(n, oracle, ctx):
    _compare(oracle, ctx, scenes.SPHERE_LUA, 1.0 / n)


@pytest.mark.parametrize("res", [0.5, 0.3, 0.11])
def test_cube_lua_fine(res, oracle, ctx):
    # exact zeros on the whole surface, side-sets on lattice planes
    _compare(oracle, ctx, scenes.CUBE_LUA, res)


@pytest.mark.parametrize("n", [9, 24, 48])
def test_xplicit_lua(n, oracle, ctx):
    _compare(oracle, ctx, scenes.XPLICIT_WITH_TOP_LUA, 15.0 / n)


def test_unit_sphere_fixture(oracle, ctx):
    # tessellate3d/src/lib.rs:486-490 (test_convert_mesh) with the reference's own UnitSphere
    t = mc.Tape()
    f = mc.LObject(t, t.unit_sphere_sq())
    mesh = mc.IsosurfaceStuffing(f, 0.1, 0.2, ctx=ctx).tessellate()
    sc = oracle.OracleScene()
    om = sc.tessellate(sc.unit_sphere_sq(), 0.1, 0.2)
    canon.assert_same_mesh(canon.canonical(mesh.vertices(), mesh.elems()), canon.canonical(om.vertices(), om.elems()))
    assert mesh.stats()["stats"] == om.stats()["stats"]


@pytest.mark.parametrize("flags", [_lib.MC_OPT_NO_TILE_PRUNING, _lib.MC_OPT_FORCE_TILE_PRUNING])
@pytest.mark.parametrize("name,res", [("cyl_cone", 0.21), ("cyl", 0.043), ("mixed", 0.031), ("xplicit", 15.0 / 40),
                                      ("sphere", 1.0 / 20), ("cube", 0.11)])
def test_ragged_scenes(name, res, flags, oracle, ctx):
    # non-cubic lattices, rotated / scaled / translated objects, cone + hessian planes; with the
    # plain evaluator and with the tile-specialised one (mc_prune.cuh)
    _compare(oracle, ctx, SCRIPTS[name], res, flags=flags)


@pytest.mark.gpu
@pytest.mark.parametrize("scene,n", [("csg", 90), ("sphere", 61), ("xplicit", 48)])
def test_plain_lattice_quality_paths_agree(scene, n, ctx):
    """check_for_inverted_tets over the plain lattice tets: one evaluation per global (step, parity)
    class vs. the per-tile classes (MC_OPT_PER_TILE_QUALITY) -- same aspect-ratio range, same count."""
    if scene == "csg":
        g = mc.Geom()
        obj, res = scenes.synthetic_csg(g, 20, seed=3), 1.0 / n
    elif scene == "sphere":
        obj, res = luascad.eval(scenes.SPHERE_LUA)["sphere"].obj, 1.0 / n
    else:
        obj, res = luascad.eval(scenes.XPLICIT_LUA)["xplicit"].obj, 15.0 / n
    obj.backend.set_parameters(0.1, 1.0)
    a = mc.IsosurfaceStuffing(obj, res, 0.2, ctx=ctx, allow_diverged=True).tessellate().stats()
    b = mc.IsosurfaceStuffing(obj, res, 0.2, ctx=ctx, flags=_lib.MC_OPT_PER_TILE_QUALITY,
                              allow_diverged=True).tessellate().stats()
    for key in ("ntets_stage", "stats", "aspect_ratio", "inverted", "n_tets"):
        assert a[key] == b[key], key
    assert a["aspect_ratio"][0] <= a["aspect_ratio"][1] <= 1.0 + 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("scene,n", [("csg", 90), ("sphere", 61), ("xplicit", 48), ("cube", 33)])
def test_shape_bytes_reproduce_the_classification(scene, n, ctx):
    """The tet emission rebuilds the pieces of a trimmed tet from the one-byte shape k_classify_cells
    recorded (trim case x sorted node order, mc_shapes.hpp).  MC_OPT_RECLASSIFY_EMIT makes it read the
    field and run the 4-sort again: same tets, in the same places."""
    if scene == "csg":
        g = mc.Geom()
        obj, res = scenes.synthetic_csg(g, 20, seed=3), 1.0 / n
    elif scene == "sphere":
        obj, res = luascad.eval(scenes.SPHERE_LUA)["sphere"].obj, 1.0 / n
    elif scene == "cube":
        obj, res = luascad.eval(scenes.CUBE_LUA)["cube"].obj, 1.0 / n
    else:
        obj, res = luascad.eval(scenes.XPLICIT_LUA)["xplicit"].obj, 15.0 / n
    obj.backend.set_parameters(0.1, 1.0)
    a = mc.IsosurfaceStuffing(obj, res, 0.2, ctx=ctx, allow_diverged=True).tessellate()
    b = mc.IsosurfaceStuffing(obj, res, 0.2, ctx=ctx, flags=_lib.MC_OPT_RECLASSIFY_EMIT, allow_diverged=True).tessellate()
    assert np.array_equal(a.elems(), b.elems())
    assert np.array_equal(a.vertices().view(np.uint64), b.vertices().view(np.uint64))
    assert a.stats()["stats"] == b.stats()["stats"]
    assert sorted(a.boundary().sides()) == sorted(b.boundary().sides())


@pytest.mark.parametrize("nprim,n", [(64, 72), (24, 100)])
def test_tile_pruning_is_value_neutral(nprim, n, ctx):
    """The tile-specialised evaluator must produce the same bits as the full program at every
    lattice point (and prune most of the program)."""
    g = mc.Geom()
    root = scenes.synthetic_csg(g, nprim, seed=nprim)
    g.backend.set_parameters(0.1, 1.0)
    res = 1.0 / n
    a = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx, flags=_lib.MC_OPT_NO_TILE_PRUNING).tessellate()
    # KEEP_PHI: evaluate every tile (by default tiles proven to lie deep inside / outside get a
    # placeholder of the right sign instead of values nobody reads)
    b = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx,
                              flags=_lib.MC_OPT_FORCE_TILE_PRUNING | _lib.MC_OPT_KEEP_PHI).tessellate()
    c = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx, flags=_lib.MC_OPT_FORCE_TILE_PRUNING).tessellate()
    assert np.array_equal(a.vertices(), c.vertices()) and np.array_equal(a.elems(), c.elems())
    with pytest.raises(_lib.McError) as ei:  # placeholders where the sign was proven: not the field
        c.phi()
    assert ei.value.code == _lib.MC_ERR_STATE
    pa, pb = a.phi(), b.phi()
    assert np.array_equal(pa.view(np.uint64), pb.view(np.uint64))
    assert np.array_equal(a.vertices(), b.vertices()) and np.array_equal(a.elems(), b.elems())
    pr = b.stats()["prune"]
    assert pr["tiles"] > 0 and pr["avg_words"] < 0.5 * pr["full_words"]
    assert a.stats()["prune"]["tiles"] == 0


def test_synthetic_csg(oracle, ctx):
    # BASELINE config 4's scene shape at a lattice the oracle finishes in seconds
    g = mc.Geom()
    root = scenes.synthetic_csg(g, 16, seed=64)
    g.backend.set_parameters(0.1, 1.0)
    res = 1.0 / 40
    mesh = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx).tessellate()
    assert mesh.stats()["prune"]["tiles"] > 0  # the tile-specialised evaluator is the default here
    go = mc.Geom(oracle.OracleScene())
    ro = scenes.synthetic_csg(go, 16, seed=64)
    ro.backend.set_parameters_node(ro.node, 0.1, 1.0)
    om = ro.backend.tessellate(ro.node, res, 0.2)
    canon.assert_same_mesh(canon.canonical(mesh.vertices(), mesh.elems()), canon.canonical(om.vertices(), om.elems()))
    assert mesh.stats()["stats"] == om.stats()["stats"]


def test_synthetic_csg_256_primitives(oracle, ctx):
    # BASELINE config 5's scene shape: the ~100 KB program needs the opt-in shared memory path
    g = mc.Geom()
    root = scenes.synthetic_csg(g, 256, seed=256, size_scale=0.5)
    g.backend.set_parameters(0.1, 1.0)
    res = 1.0 / 28
    mesh = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx).tessellate()
    assert mesh.stats()["prune"]["tiles"] > 0
    go = mc.Geom(oracle.OracleScene())
    ro = scenes.synthetic_csg(go, 256, seed=256, size_scale=0.5)
    ro.backend.set_parameters_node(ro.node, 0.1, 1.0)
    om = ro.backend.tessellate(ro.node, res, 0.2)
    canon.assert_same_mesh(canon.canonical(mesh.vertices(), mesh.elems()), canon.canonical(om.vertices(), om.elems()))
    assert mesh.stats()["stats"] == om.stats()["stats"]


def test_tile_program_over_the_shared_memory_budget(oracle, ctx):
    """Tiles whose specialised program does not fit the 32 KB shared-memory budget evaluate the full
    program from global memory instead (same values): 320 overlapping primitives on a lattice of
    2 x 2 x 2 tiles, where every tile sees nearly the whole scene."""
    g = mc.Geom()
    root = scenes.synthetic_csg(g, 320, seed=7)
    g.backend.set_parameters(0.1, 1.0)
    res = 1.0 / 20
    mesh = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx, flags=_lib.MC_OPT_KEEP_PHI).tessellate()
    pr = mesh.stats()["prune"]
    assert pr["tiles"] > 0 and pr["full_words"] > 4096 and pr["avg_words"] > 4096, pr
    go = mc.Geom(oracle.OracleScene())
    ro = scenes.synthetic_csg(go, 320, seed=7)
    ro.backend.set_parameters_node(ro.node, 0.1, 1.0)
    om = ro.backend.tessellate(ro.node, res, 0.2)
    canon.assert_same_mesh(canon.canonical(mesh.vertices(), mesh.elems()), canon.canonical(om.vertices(), om.elems()))
    assert mesh.stats()["stats"] == om.stats()["stats"]
    # bit-identical field against the un-specialised evaluator
    ref = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx, flags=_lib.MC_OPT_NO_TILE_PRUNING | _lib.MC_OPT_KEEP_PHI).tessellate()
    assert np.array_equal(mesh.phi().view(np.uint64), ref.phi().view(np.uint64))


def _check_sidesets_against_point_eval(mesh, out, ctx):
    """Side-set membership recomputed from the boundary object's values at the surface nodes
    (mc_eval_points, a different kernel path) must reproduce the side-sets (src/main.rs:87-106)."""
    eps = np.finfo(np.float64).eps
    v = mesh.vertices()
    sides = mesh.boundary().array().astype(np.int64)
    e = mesh.elems()
    tri = e[sides[:, 0][:, None], canon.FACE[sides[:, 1]]].astype(np.int64)
    nodes, inv = np.unique(tri, return_inverse=True)
    sizes = {}
    for ss in mesh.side_sets():
        obj = out[ss.name()].boundary()
        val = mc.ObjectAdaptor(obj, eps, ctx).value(v[nodes], slack=eps)
        ok = ~((val > eps) | (val < -eps))
        member = ok[inv.reshape(tri.shape)].all(axis=1)
        got = ss.array().astype(np.int64)
        want = sides[member]
        assert got.shape == want.shape and np.array_equal(got, want), ss.name()
        sizes[ss.name()] = len(got)
    # surface is closed (even edge counts)
    edges = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(edges, axis=0, return_counts=True)
    assert (cnt % 2 == 0).all()
    return sizes


def test_full_size_xplicit_512_sideset(ctx):
    """BASELINE config 3: xplicit.lua at 512^3 with an added :boundary() object, checked through
    properties that do not need the oracle (which cannot finish this size).  Run with
    --tessellation-error 0.8: with the default 0.2 the reference's cut_edge never terminates on
    this scene.  The side-set tolerance of the reference is f64::EPSILON,
    so after scale(15) (1/15 is not representable) the set may well be empty; what is checked is
    that it is exactly what the boundary object's values say."""
    script = scenes.XPLICIT_WITH_TOP_LUA
    mesh = main.run(script, tessellation_resolution=15.0 / 512, tessellation_error=0.8, ctx=ctx)
    st = mesh.stats()
    assert st["dim"] == [512, 512, 512] and st["inverted"] == 0
    _check_sidesets_against_point_eval(mesh, luascad.eval(script), ctx)


def test_full_size_cube_256_sidesets(ctx):
    """examples/cube.lua on a 256^3 lattice: every value is exact.  The `top` object is the cube
    intersected with a 0.01 thick slab around z = 0.5, so its zero set holds the 2 x 256^2 triangles
    of the plane z = 0.5 plus the top row (z >= 0.5 - 1/256 > 0.495) of the four side walls."""
    mesh = main.run(scenes.CUBE_LUA, tessellation_resolution=1.0 / 256, ctx=ctx)
    st = mesh.stats()
    assert st["dim"] == [256, 256, 256] and st["n_tets"] == 5 * 256 ** 3 and st["n_vertices"] == 257 ** 3
    sizes = _check_sidesets_against_point_eval(mesh, luascad.eval(scenes.CUBE_LUA), ctx)
    expect = 2 * 256 * 256 + 4 * 256 * 2
    assert sizes == {"top": expect, "bottom": expect}
    assert len(mesh.boundary()) == 12 * 256 * 256


def test_tet_quality_matches_oracle(oracle, ctx):
    rng = np.random.default_rng(3)
    coords = rng.normal(size=(400, 3))
    tets = rng.integers(0, 400, size=(1000, 4)).astype(np.uint64)
    vol, ar = ctx.tet_quality(coords, tets)
    om = oracle.OracleMesh.from_arrays(coords, tets)
    want_v = np.array([om.volume(i) for i in range(len(tets))])
    want_a = np.array([om.aspect_ratio(i) for i in range(len(tets))])
    assert np.array_equal(vol, want_v)
    assert np.array_equal(ar, want_a, equal_nan=True)
    # mesh/src/lib.rs:817-826: the regular tet has aspect ratio exactly 1
    _, ar = ctx.tet_quality([(1, 1, 1), (-1, -1, 1), (-1, 1, -1), (1, -1, -1)], [[0, 1, 3, 2]])
    assert ar[0] == 1.0


def test_error_behaviour(ctx):
    # infinite bounding box: the reference panics in to_usize().unwrap() (tessellate3d/src/lib.rs:44-46)
    g = mc.Geom()
    with pytest.raises(_lib.McError) as e:
        mc.IsosurfaceStuffing(g.plane_x(1.0), 0.1, 0.2, ctx=ctx).tessellate()
    assert e.value.code == _lib.MC_ERR_BBOX
    with pytest.raises(_lib.McError) as e:
        mc.IsosurfaceStuffing(g.sphere(1.0), -1.0, 0.2, ctx=ctx).tessellate()
    assert e.value.code == _lib.MC_ERR_ARG


def test_empty_result(oracle, ctx):
    # a volume with no lattice point inside: zero tets, zero vertices
    script = "return { s = geom.sphere(0.01):translate(0.05, 0.05, 0.05):intersection(geom.cube(1, 1, 1, 0), 0):volume() }"
    mesh, om = _compare(oracle, ctx, script, 0.5)
    assert mesh.num_elems() == len(om.elems())


def test_full_size_properties_sphere_256(ctx):
    """BASELINE config 2 at full size (256^3), checked through size-independent properties."""
    mesh = main.run(scenes.SPHERE_LUA, tessellation_resolution=1.0 / 256, ctx=ctx)
    st = mesh.stats()
    assert st["dim"] == [256, 256, 256]
    assert st["inverted"] == 0
    v, e = mesh.vertices(), mesh.elems().astype(np.int64)
    assert e.max() == len(v) - 1 and np.unique(e).size == len(v)       # compact: every vertex used
    a, b, c, d = v[e[:, 0]], v[e[:, 1]], v[e[:, 2]], v[e[:, 3]]
    vol = np.einsum("ij,ij->i", a - d, np.cross(b - d, c - d)) / 6.0
    assert (vol > 0).all()
    assert abs(vol.sum() - math.pi / 6) / (math.pi / 6) < 2e-4              # 4/3 pi 0.5^3
    # the surface is closed: every edge of the boundary triangles is shared by an even number of
    # them (the reference's stuffing itself produces a few non-manifold edges with 4, see the
    # oracle at N = 32..128)
    sides = np.array(sorted(mesh.boundary().sides()), dtype=np.int64)
    tri = e[sides[:, 0][:, None], canon.FACE[sides[:, 1]]]
    edges = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(edges, axis=0, return_counts=True)
    assert (cnt % 2 == 0).all() and (cnt == 2).mean() > 0.97
    # boundary vertices are bisected cut points (within 0.2 res of the surface) or warped lattice
    # vertices (moved along an edge by linear interpolation): all well within one cell
    r = np.linalg.norm(v[np.unique(tri)], axis=1)
    assert np.abs(r - 0.5).max() <= 1.0 / 256
    # idempotence: a second run gives the identical arrays
    mesh2 = main.run(scenes.SPHERE_LUA, tessellation_resolution=1.0 / 256, ctx=ctx)
    assert np.array_equal(mesh2.vertices(), v) and np.array_equal(mesh2.elems().astype(np.int64), e)


@pytest.mark.parametrize("script,res", [(scenes.CUBE_LUA, 1.0 / 9), (scenes.XPLICIT_WITH_TOP_LUA, 15.0 / 70)])
def test_exodus_layout(script, res, ctx):
    """exodus::write's arrays (exodus/src/lib.rs:55-63 coord, :84-96 connect1, :118-132 elem_ss / side_ss)
    produced on the device against a numpy restatement of the writer's loops."""
    mesh = main.run(script, tessellation_resolution=res, ctx=ctx)
    ex = mesh.exodus_arrays()
    v, e = mesh.vertices(), mesh.elems()
    assert ex["coord"].shape == (3, len(v)) and np.array_equal(ex["coord"], v.T)          # vertices[i + num_nodes * j] = v[j]
    assert ex["connect1"].dtype == np.int64 and np.array_equal(ex["connect1"], e.astype(np.int64) + 1)   # e.node(j) as i64 + 1
    assert len(ex["side_sets"]) == len(mesh.side_sets())
    for (elem_ss, side_ss), ss in zip(ex["side_sets"], mesh.side_sets()):
        a = ss.array().astype(np.int64)
        assert elem_ss.dtype == np.int32 and np.array_equal(elem_ss, a[:, 0] + 1) and np.array_equal(side_ss, a[:, 1] + 1)


def _canonical_tris(coords, tris):
    """Triangles as coordinate triples, rotated (cyclic, orientation preserved) to start at their smallest
    vertex, sorted."""
    c = np.asarray(coords, dtype=np.float64)[np.asarray(tris, dtype=np.int64)] + 0.0    # (m, 3, 3)
    key = np.lexsort((c[..., 2], c[..., 1], c[..., 0]), axis=1)[:, 0]                   # index of the smallest vertex
    rot = (np.arange(3)[None, :] + key[:, None]) % 3
    c = c[np.arange(len(c))[:, None], rot].reshape(len(c), 9)
    return c[np.lexsort(c.T[::-1])]


@pytest.mark.parametrize("script,res", [(scenes.SPHERE_LUA, 1.0 / 23), (scenes.XPLICIT_LUA, 15.0 / 31)])
def test_to_boundary(script, res, oracle, ctx):
    """Mesh::to_boundary (mesh/src/lib.rs:359-386): one Tri3 per surface side in Tet4::face node order,
    vertices compacted — against the same construction on the oracle's mesh."""
    mesh = main.run(script, tessellation_resolution=res, ctx=ctx)
    v, t = mesh.to_boundary()
    om, _, _ = _oracle_run(oracle, script, res)
    ov, oe, osides = om.vertices(), om.elems().astype(np.int64), om.boundary().astype(np.int64)
    otris = oe[osides[:, 0][:, None], canon.FACE[osides[:, 1]]]
    assert t.shape == otris.shape
    assert len(v) == np.unique(otris).size and np.unique(t).size == len(v) and int(t.max()) == len(v) - 1   # compact
    assert np.array_equal(_canonical_tris(v, t), _canonical_tris(ov, otris))


@pytest.mark.gpu
def test_boundary_capacity_overflow_repeats_the_pass(ctx, monkeypatch):
    """Mesh::boundary writes the sides in one pass into a buffer sized from a guess; when the guess is short
    the pass only counts and is repeated with the exact size (forced here through MC_BOUNDARY_CAPACITY)."""
    obj = luascad.eval(scenes.SPHERE_LUA)["sphere"].obj
    obj.backend.set_parameters(0.1, 1.0)
    a = mc.IsosurfaceStuffing(obj, 1.0 / 40, 0.2, ctx=ctx).tessellate()
    want = a.boundary().array()
    assert len(want) > 1000
    monkeypatch.setenv("MC_BOUNDARY_CAPACITY", "100")
    b = mc.IsosurfaceStuffing(obj, 1.0 / 40, 0.2, ctx=ctx).tessellate()
    got = b.boundary().array()
    assert np.array_equal(got, want)
    key = got[:, 0] * 4 + got[:, 1]
    assert np.all(key[1:] > key[:-1])   # the device writes the sides in (elem, side) order
