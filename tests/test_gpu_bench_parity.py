"""Parity at the BENCHMARK configuration (BASELINE config 4: the 64-primitive smooth CSG scene of
numpy default_rng(64)):

  * against the oracle at lattices it finishes in seconds (full comparison: mesh, statistics,
    Mesh::boundary) — tessellate3d/src/lib.rs:420-431;
  * at 512^3 and at the full 1024^3 through a device-side checksum: the default path (interval
    sign proofs, +-1 placeholders, tile-specialised programs) against MC_OPT_NO_TILE_PRUNING (the
    full program at every lattice point), one mesh resident at a time;
  * topology against the oracle on the PLATFORM libm (what the Rust binary executes): the
    number of differing vertices / tets is reported and bounded.
"""
import json
import os

import numpy as np
import pytest

import moosecad_b200 as mc
from moosecad_b200 import _lib, luascad, scenes

import canon

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_scene(backend=None, nprim=64):
    g = mc.Geom(backend) if backend is not None else mc.Geom()
    root = scenes.synthetic_csg(g, nprim, seed=nprim)
    if backend is None:
        g.backend.set_parameters(0.1, 1.0)
    else:
        root.backend.set_parameters_node(root.node, 0.1, 1.0)
    return g, root


@pytest.mark.parametrize("n", [40, 64])
def test_bench_scene_vs_oracle(n, oracle, ctx):
    res = 1.0 / n
    g, root = _bench_scene()
    mesh = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx).tessellate()
    st = mesh.stats()
    assert st["prune"]["tiles"] > 0          # the default (tile-specialised, sign-proving) path ran
    go, ro = _bench_scene(oracle.OracleScene())
    om = ro.backend.tessellate(ro.node, res, 0.2)
    assert om.ok()
    ost = om.stats()
    a = canon.canonical(mesh.vertices(), mesh.elems())
    b = canon.canonical(om.vertices(), om.elems())
    canon.assert_same_mesh(a, b, bitwise=True)
    assert st["dim"] == ost["dim"] == [n, n, n]
    assert st["ntets_stage"] == ost["ntets_stage"]
    assert st["stats"] == ost["stats"]
    assert st["n_warped"] == ost["n_warped"] and st["n_cut"] == ost["n_cut"]
    assert st["inverted"] == ost["inverted"]
    assert ost["n_exterior_skipped_by_swap"] == 0
    np.testing.assert_allclose(st["aspect_ratio"], ost["aspect_ratio"], rtol=1e-12)
    gb = canon.canonical(mesh.vertices(), mesh.elems(), {"s": sorted(mesh.boundary().sides())})[2]["s"]
    ob = canon.canonical(om.vertices(), om.elems(), {"s": om.boundary()})[2]["s"]
    assert np.array_equal(gb, ob)
    # the device checksum is the numpy mirror applied to the oracle's mesh
    cs = mesh.checksum()
    assert (cs["vertices"], cs["tets"]) == canon.mesh_checksum(om.vertices(), om.elems())
    assert cs["tets_skipped"] == 0


def test_checksum_sees_a_single_flipped_tet(ctx):
    g, root = _bench_scene(nprim=8)
    mesh = mc.IsosurfaceStuffing(root, 1.0 / 24, 0.2, ctx=ctx).tessellate()
    v, e = mesh.vertices(), mesh.elems().copy()
    cs = mesh.checksum()
    assert (cs["vertices"], cs["tets"]) == canon.mesh_checksum(v, e)
    perm = np.random.default_rng(0).permutation(len(v))     # renumbering + reordering: same checksum
    v2 = np.empty_like(v)
    v2[perm] = v
    e2 = perm[e.astype(np.int64)][::-1][:, [2, 0, 1, 3]]
    assert canon.mesh_checksum(v2, e2) == (cs["vertices"], cs["tets"])
    e[len(e) // 2, [0, 1]] = e[len(e) // 2, [1, 0]]          # one inverted tet: different
    assert canon.mesh_checksum(v, e)[1] != cs["tets"]


def _run_and_summarise(root, res, ctx, flags):
    mesh = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx, flags=flags).tessellate()
    st = mesh.stats()
    out = {"checksum": mesh.checksum(), "stats": st["stats"], "ntets_stage": st["ntets_stage"],
           "n_vertices": st["n_vertices"], "n_tets": st["n_tets"], "n_warped": st["n_warped"], "n_cut": st["n_cut"],
           "inverted": st["inverted"], "aspect_ratio": st["aspect_ratio"], "prune_tiles": st["prune"]["tiles"],
           "sdf_skipped_tiles": st["tiles"]["sdf_skipped"]}
    del mesh
    return out


@pytest.mark.parametrize("n", [512, 1024])
def test_pruning_ab_at_bench_size(n, ctx):
    """Sign proofs / placeholders / tile programs change nothing: same mesh as evaluating the
    whole program at every one of the (n+1)^3 lattice points."""
    g, root = _bench_scene()
    res = 1.0 / n
    a = _run_and_summarise(root, res, ctx, 0)
    b = _run_and_summarise(root, res, ctx, _lib.MC_OPT_NO_TILE_PRUNING)
    ctx.trim()
    assert a["prune_tiles"] > 0 and a["sdf_skipped_tiles"] > 0 and b["prune_tiles"] == 0
    assert a["checksum"]["tets_skipped"] == 0
    for k in ("checksum", "stats", "ntets_stage", "n_vertices", "n_tets", "n_warped", "n_cut", "inverted", "aspect_ratio"):
        assert a[k] == b[k], k
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"r02_pruning_ab_{n}.json"), "w") as f:
            json.dump({"default": a, "no_tile_pruning": b}, f, indent=1)


@pytest.mark.parametrize("scene,n", [("xplicit", 48), ("csg64", 48)])
def test_topology_vs_platform_libm(scene, n, oracle, ctx):
    """The reference calls the platform libm for exp / ln; the device (and oracle libm_mode 1) use a
    portable restatement that differs from glibc by < 1 ulp.  A 1-ulp difference in phi can flip
    `alpha < 0.3`, the 4-sort or the `close` test: count what actually differs."""
    if scene == "xplicit":
        res = 15.0 / n
        root = luascad.eval(scenes.XPLICIT_LUA)["xplicit"].obj
        root.backend.set_parameters(0.1, 1.0)
        ro = luascad.eval(scenes.XPLICIT_LUA, backend=oracle.OracleScene())["xplicit"].obj
        ro.backend.set_parameters_node(ro.node, 0.1, 1.0)
    else:
        res = 1.0 / n
        _, root = _bench_scene()
        _, ro = _bench_scene(oracle.OracleScene())
    mesh = mc.IsosurfaceStuffing(root, res, 0.2, ctx=ctx).tessellate()
    oracle.set_libm_mode(0)
    try:
        om = ro.backend.tessellate(ro.node, res, 0.2)
    finally:
        oracle.set_libm_mode(1)
    st, ost = mesh.stats(), om.stats()
    d = canon.mesh_diff(mesh.vertices(), mesh.elems(), om.vertices(), om.elems())
    d.update({"scene": scene, "n": n, "stats_gpu": st["stats"], "stats_oracle_libm0": ost["stats"],
              "n_warped": [st["n_warped"], ost["n_warped"]], "n_cut": [st["n_cut"], ost["n_cut"]]})
    print("libm topology diff:", json.dumps(d))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"r02_libm_topology_diff_{scene}_{n}.json"), "w") as f:
            json.dump(d, f, indent=1)
    # Identical on the CSG scene.  xplicit.lua is mirror-symmetric in y / z: a vertex whose two best
    # warp candidates are equal in exact arithmetic ("first visit wins" decides) sees them differ by
    # an ulp under one libm and not the other, and is pulled along the mirror-image edge: 8 of 27 288
    # vertices (one per octant) at N = 48, with the ~23 tets around each.  Counts and statistics agree.
    assert st["stats"] == ost["stats"] and st["n_warped"] == ost["n_warped"] and st["n_cut"] == ost["n_cut"]
    assert d["vertices_a"] == d["vertices_b"] and d["tets_a"] == d["tets_b"]
    assert d["vertices_only_a"] <= max(8, d["vertices_a"] // 2000), d
    assert d["tets_only_a"] <= max(200, d["tets_a"] // 400), d
    if scene == "csg64":
        assert d["tets_only_a"] == 0 and d["vertices_only_a"] == 0, d
