"""Pins the oracle (oracle/oracle.cpp) to every known answer the reference's own tests hold
for the hot path.  CPU only."""
import numpy as np
import pytest

from moosecad_b200 import luascad, scenes


def test_cut_edge(oracle):
    # tessellate3d/src/lib.rs:492-507
    sc = oracle.OracleScene()
    f = sc.unit_sphere_sq()
    idx, pos, phi_new = sc.cut_edge_test(f, 0.1, 0.0001, (-1.5, 0.0, 0.0), (-0.4, 0.0, 0.0))
    assert idx == 2
    assert phi_new == 0.0
    assert np.linalg.norm(np.array(pos) - np.array([-1.0, 0.0, 0.0])) < 0.00001


def test_convert_mesh_terminates(oracle):
    # tessellate3d/src/lib.rs:486-490: UnitSphere at res 0.1 / err 0.2 must terminate
    sc = oracle.OracleScene()
    m = sc.tessellate(sc.unit_sphere_sq(), 0.1, 0.2)
    assert m.ok()
    st = m.stats()
    assert st["dim"] == [20, 20, 20]
    assert len(m.elems()) > 0 and st["inverted"] == 0


TILE_COORDS = [(-.5, -.5, -.5), (.5, -.5, -.5), (.5, .5, -.5), (-.5, .5, -.5),
               (-.5, -.5, .5), (.5, -.5, .5), (.5, .5, .5), (-.5, .5, .5)]
TILE_TETS = [[0, 1, 5, 2], [0, 2, 5, 7], [0, 7, 5, 4], [2, 6, 5, 7], [0, 2, 7, 3]]


def test_tile_volume(oracle):
    # mesh/src/lib.rs:747-777
    m = oracle.OracleMesh.from_arrays(TILE_COORDS, TILE_TETS)
    assert len(m.boundary()) == 12
    total = m.volume(0)
    for i in range(1, 5):
        total = total + m.volume(i)  # left to right like the reference (python's sum() compensates)
    assert total < 1.0
    assert total > 0.999999
    assert all(m.volume(i) > 0 for i in range(5))


def test_tile_tables_match_reference(oracle):
    # tessellate3d/src/lib.rs:436-454
    L = oracle.lib()
    lat = np.ctypeslib.as_array(L.orc_tile_lattice(), shape=(8, 3))
    tets = np.ctypeslib.as_array(L.orc_tile_tets(), shape=(5, 4))
    assert lat.tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]]
    assert tets.tolist() == TILE_TETS


def test_bcc_tet4_boundary(oracle):
    # mesh/src/lib.rs:727-745
    m = oracle.OracleMesh.from_arrays([(.5, .5, .5), (1.5, .5, .5), (1., 0., 1.), (1., 1., 1.)], [[0, 1, 2, 3]])
    assert len(m.boundary()) == 4
    s = m.face_normal(0, 0) + m.face_normal(0, 1) + m.face_normal(0, 2) + m.face_normal(0, 3)
    assert s.tolist() == [0.0, 0.0, 0.0]


def test_compact_mesh(oracle):
    # mesh/src/lib.rs:779-795
    coords = [(0, 0, 0), (.5, .5, .5), (0, 0, 0), (1.5, .5, .5), (0, 0, 0), (1, 0, 1), (0, 0, 0), (1, 1, 1), (0, 0, 0)]
    m = oracle.OracleMesh.from_arrays(coords, [[1, 3, 5, 7]])
    m.compact()
    assert len(m.vertices()) == 4
    assert m.elems().tolist() == [[0, 1, 2, 3]]


def test_aspect_ratio(oracle):
    # mesh/src/lib.rs:817-826
    m = oracle.OracleMesh.from_arrays([(1, 1, 1), (-1, -1, 1), (-1, 1, -1), (1, -1, -1)], [[0, 1, 3, 2]])
    assert m.aspect_ratio(0) == 1.0


def test_luascad_primitives(oracle):
    # luascad/src/lib.rs:339-378
    res = luascad.eval("return { cube = geom.cube(1, 1, 1, 0.0):translate(1.0, 1.0, 1.0):volume() }",
                       backend=oracle.OracleScene())
    assert "cube" in res
    res = luascad.eval("return { sphere = geom.sphere(0.5):volume() }", backend=oracle.OracleScene())
    assert "sphere" in res
    res = luascad.eval("""
            cube = geom.cube(1, 1, 1, 0.3)
            sphere = geom.sphere(0.5)
            diff = cube:difference(sphere, 0.3)
            union = cube:union(sphere, 0.3)
            return { xplicit = diff:scale(15, 15, 15):volume() }
        """, backend=oracle.OracleScene())
    assert "xplicit" in res
    res = luascad.eval(scenes.CUBE_LUA, backend=oracle.OracleScene())
    assert set(res) == {"cube", "top", "bottom"}
    top = res["top"].boundary()
    assert top.backend.approx_value(top.node, (0.0, 0.0, 0.5), 0.0) == 0.0
    assert top.backend.approx_value(top.node, (0.0, 0.0, 1.0), 0.0) == 0.5


def test_config1_cube_lua(oracle):
    # BASELINE config 1: examples/cube.lua at --tessellation-resolution 1.0
    res = luascad.eval(scenes.CUBE_LUA, backend=oracle.OracleScene())
    for o in res.values():
        o.obj.backend.set_parameters_node(o.obj.node, 0.1, 1.0)
    cube = res["cube"].volume()
    m = cube.backend.tessellate(cube.node, 1.0, 0.2)
    st = m.stats()
    assert st["dim"] == [1, 1, 1]
    assert len(m.vertices()) == 8 and len(m.elems()) == 5
    assert len(m.boundary()) == 12
    v, e = m.vertices(), m.elems()
    for name, z in (("top", 0.5), ("bottom", -0.5)):
        b = res[name].boundary()
        ss = m.add_sideset(b.backend, b.node)
        assert len(ss) == 2
        for elem, side in ss.tolist():
            tri = [e[elem][k] for k in ([0, 1, 2], [0, 2, 3], [0, 3, 1], [1, 3, 2])[side]]
            assert all(v[n][2] == z for n in tri)
    # first-touch vertex order L0,L1,L5,L2,L7,L4,L6,L3
    assert v.tolist() == [[-.5, -.5, -.5], [.5, -.5, -.5], [.5, -.5, .5], [.5, .5, -.5],
                          [-.5, .5, .5], [-.5, -.5, .5], [.5, .5, .5], [-.5, .5, -.5]]


def test_xplicit_bbox_and_dim(oracle):
    # scale(15) inverse gives exactly +-7.5 => N = 512 at res 15/512
    res = luascad.eval(scenes.XPLICIT_LUA, backend=oracle.OracleScene())
    o = res["xplicit"].volume()
    assert o.bbox() == [-7.5, -7.5, -7.5, 7.5, 7.5, 7.5]


def test_sphere_volume_sanity(oracle):
    # sum of signed tet volumes ~ 4/3 pi r^3 (SURVEY.md §8c probes); coarse lattice => loose bound
    res = luascad.eval(scenes.SPHERE_LUA, backend=oracle.OracleScene())
    s = res["sphere"].volume()
    m = s.backend.tessellate(s.node, 1.0 / 32, 0.2)
    st = m.stats()
    assert st["inverted"] == 0 and st["n_exterior_skipped_by_swap"] == 0
    raw = oracle.OracleMesh.from_arrays(m.vertices(), m.elems())
    vol = sum(raw.volume(i) for i in range(len(m.elems())))
    assert abs(vol - 4.0 / 3.0 * np.pi * 0.125) / (4.0 / 3.0 * np.pi * 0.125) < 0.02


@pytest.mark.parametrize("mode", [0, 1])
def test_libm_modes_agree(oracle, mode):
    # portable exp/ln vs the platform libm: < 1 ulp apart on the smooth-boolean argument range
    oracle.set_libm_mode(mode)
    try:
        xs = np.concatenate([np.linspace(-40, 5, 2001), np.array([0.0, 1e-10, -1e-10, 700.0, -700.0])])
        L = oracle.lib()
        got = np.array([L.orc_exp(float(x)) for x in xs])
        want = np.exp(xs)
        ulp = np.abs(got - want) / np.spacing(want)
        assert ulp.max() <= 1.0
        ys = np.concatenate([np.exp(np.linspace(-30, 30, 2001)), np.array([1.0, 1.0 + 1e-15, 1.0 - 1e-15, 0.5, 2.0])])
        got = np.array([L.orc_ln(float(y)) for y in ys])
        want = np.log(ys)
        ulp = np.abs(got - want) / np.maximum(np.spacing(np.abs(want)), 5e-324)
        assert ulp.max() <= 1.0
    finally:
        oracle.set_libm_mode(1)


def test_named_assumptions_switch(oracle):
    """The recalled implicit3d / bbox / nalgebra details are named switches of the
    oracle.  The reference-held known answers (luascad test_primitives, config 1) hold under every
    alternative - they do not discriminate them - while a rotated smooth scene does (see
    profiles/r02_assumption_sensitivity.json, scripts/assumption_sensitivity.py)."""
    def cube_answers():
        res = luascad.eval(scenes.CUBE_LUA, backend=oracle.OracleScene())
        top = res["top"].boundary()
        return (top.backend.approx_value(top.node, (0.0, 0.0, 0.5), 0.0),
                top.backend.approx_value(top.node, (0.0, 0.0, 1.0), 0.0))

    def rotated_value():
        sc = oracle.OracleScene()
        g = luascad.Geom(sc)
        o = g.cube(1, 0.6, 0.4, 0.05).rotate(0.3, 0.5, 0.7).union(g.sphere(0.35).translate(0.2, 0.1, 0), 0.1)
        return o.backend.approx_value(o.node, (0.31, -0.22, 0.17), 0.0)

    try:
        oracle.reset_assumptions()
        base = rotated_value()
        assert cube_answers() == (0.0, 0.5)
        changed = set()
        for name, (_, values) in oracle.ASSUMPTIONS.items():
            for v in values:
                if v == 0:
                    continue
                oracle.reset_assumptions()
                assert oracle.set_assumption(name, v) == 0
                assert cube_answers() == (0.0, 0.5)
                if rotated_value() != base:
                    changed.add(name)
        assert "euler_order" in changed
        with pytest.raises(ValueError):
            oracle.set_assumption("euler_order", 7)
    finally:
        oracle.reset_assumptions()
    assert rotated_value() == base
