"""CPU tests: scene lowering (host C++ behind the C ABI), the Lua-subset front-end, the C ABI's
symbol coverage, and loud failure without a GPU.  No compute calls here."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import moosecad_b200 as mc
from moosecad_b200 import _lib, luascad, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "moosecad_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 40
    handle = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/moosecad_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().mc_abi_version() == 1


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "moosecad_b200.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    assert "at::" not in hdr and "cudaStream_t" not in re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)


def _both(script, oracle):
    a = luascad.eval(script)
    b = luascad.eval(script, backend=oracle.OracleScene())
    return a, b


@pytest.mark.parametrize("script", [scenes.CUBE_LUA, scenes.SPHERE_LUA, scenes.XPLICIT_LUA, scenes.XPLICIT_WITH_TOP_LUA,
                                    "return { c = geom.cylinder(2.0, 0.5, 0.25, 0.1):rotate(0.3, 0.2, 0.1):translate(1, 2, 3):scale(2, 3, 4):volume() }",
                                    "return { c = geom.cylinder(2.0, 0.5, 0.5, 0.0):translate(0.1, 0.2, 0.3):volume() }",
                                    "a = geom.sphere(0.3):translate(0.2, 0, 0)\nb = geom.cube(0.5, 0.4, 0.3, 0.05):rotate(1, 2, 3)\nreturn { u = a:union(b, 0.1):volume(), d = a:difference(b, 0.02):boundary() }"])
def test_bbox_propagation_matches_oracle(script, oracle):
    a, b = _both(script, oracle)
    assert set(a) == set(b)
    for k in a:
        assert a[k].kind == b[k].kind
        ba, bb = a[k].obj.bbox(), b[k].obj.bbox()
        assert np.array_equal(np.array(ba), np.array(bb)), (k, ba, bb)


def test_synthetic_scene_bbox_and_program(oracle):
    g = mc.Geom()
    root = scenes.synthetic_csg(g, 64)
    go = mc.Geom(oracle.OracleScene())
    root_o = scenes.synthetic_csg(go, 64)
    assert root.bbox() == root_o.bbox() == [-0.5, -0.5, -0.5, 0.5, 0.5, 0.5]
    info = g.backend.program_info(root.node, 1.0 / 1024)
    assert info["value_depth"] <= 8 and info["point_depth"] == 1 and info["words"] < 8192


def test_lua_subset():
    out = luascad.eval("""
        -- comment
        local r = 2 * (0.25 + 1/4) - 0.5   -- 0.5
        s = geom.sphere(r)
        t = s:translate(-1, math.pi / 2, 3e-1);
        return { a = s:volume(), b = t:boundary(), }
    """)
    assert out["a"].kind == "Volume" and out["b"].kind == "Boundary"
    assert out["a"].obj.bbox() == [-0.5] * 3 + [0.5] * 3
    np.testing.assert_allclose(out["b"].obj.bbox(), [-1.5, math.pi / 2 - .5, -.2, -.5, math.pi / 2 + .5, .8])
    with pytest.raises(luascad.LuaError):
        luascad.eval("return { a = undefined_thing:volume() }")
    with pytest.raises(luascad.LuaError):
        luascad.eval("return 3")
    with pytest.raises(luascad.LuaError):
        luascad.eval("x = = 2")


def test_unsupported_and_bad_arguments():
    g = mc.Geom()
    s = g.sphere(1.0)
    with pytest.raises(_lib.McError) as e:
        s.bend(1.0)
    assert e.value.code == _lib.MC_ERR_UNSUPPORTED
    with pytest.raises(_lib.McError) as e:
        s.twist(1.0)
    assert e.value.code == _lib.MC_ERR_UNSUPPORTED
    with pytest.raises(_lib.McError) as e:
        s.scale(math.inf, 1.0, 1.0)  # 1/inf = 0 -> singular matrix; the reference panics
    assert e.value.code in (_lib.MC_ERR_SINGULAR,)
    lib = _lib.lib()
    assert lib.mc_tape_sphere(None, 1.0) == _lib.MC_ERR_ARG
    assert lib.mc_tape_translate(g.backend.handle, 9999, 0.0, 0.0, 0.0) == _lib.MC_ERR_ARG
    assert b"mc_tape_translate" in lib.mc_last_error()


def test_union_of_one_returns_child_and_transform_chain_is_one_node():
    t = mc.Tape()
    s = t.sphere(0.5)
    u = t.union([s], 0.1)
    assert t.bbox(u) == t.bbox(s)
    a = t.translate(s, 1.0, 0.0, 0.0)
    n0 = _lib.lib().mc_tape_num_nodes(t.handle)
    b = t.scale(t.rotate(a, 0.1, 0.2, 0.3), 2.0, 2.0, 2.0)
    # chained transforms compose into one affine node each (no nesting): point depth stays 1
    assert t.program_info(b, 0.1)["point_depth"] == 1
    assert _lib.lib().mc_tape_num_nodes(t.handle) == n0 + 2


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails loudly."""
    lib = _lib.lib()
    h = lib.mc_ctx_new(0, None)
    if h:
        lib.mc_ctx_free(h)
        pytest.skip("a CUDA device is present")
    assert b"no CPU fallback" in lib.mc_last_error()
    with pytest.raises(_lib.McError):
        mc.Context(0)


def test_shape_tables_are_self_consistent():
    """The tables behind the tet emission and Mesh::boundary (csrc/mc_shapes.hpp): 146 shapes (nothing, whole, 6 trim
    cases x 24 node orders), the lattice tet across every lattice face and what the two sides agree on."""
    import ctypes as C
    import numpy as np
    lib = _lib.lib()
    n = lib.mc_shape_tables(None, 0)
    assert n > 0
    buf = (C.c_uint8 * n)()
    assert lib.mc_shape_tables(buf, n) == 0
    raw = np.frombuffer(buf, dtype=np.uint8)
    NS = 2 + 6 * 24
    off = 0

    def take(dtype, shape):
        nonlocal off
        cnt = int(np.prod(shape))
        a = raw[off:off + cnt * np.dtype(dtype).itemsize].view(dtype).reshape(shape)
        off += cnt * np.dtype(dtype).itemsize
        return a

    nodes = take(np.uint64, (NS,))
    own = take(np.uint32, (NS,))
    pfk = take(np.uint16, (NS, 4))
    pfsig = take(np.uint8, (NS, 6, 12))
    fsig = take(np.uint32, (NS, 4, 6))
    geo = take(np.uint32, (8, 5, 4))
    used = take(np.uint8, (NS,))
    # pieces: case 1, 4, 6 -> one tet, 5 -> two, 2 and 3 -> three (C_CASE_N); nothing -> none
    npieces = (nodes >> np.uint64(60)).astype(int)
    assert npieces[0] == 0 and npieces[1] == 1
    for kase, want in zip(range(1, 7), (1, 3, 3, 1, 2, 1)):
        assert set(npieces[2 + 24 * (kase - 1):2 + 24 * kase]) == {want}
    assert np.all(nodes[2:] != 0) and len(set(nodes[2:].tolist())) >= 100   # (some orders of a case give the same pieces)
    assert np.all((own >> 24) == 4 * npieces)
    # the untouched lattice tet: side s lies in the lattice face opposite node 3, 1, 2, 0 (Tet4::face)
    assert pfk[1].tolist() == [1 << 3, 1 << 1, 1 << 2, 1 << 0] and own[1] & 0xffffff == 0 and used[1] == 0xf
    assert np.all(fsig[1] == 0x16)                      # a whole lattice face, whatever the rank order
    # every piece-face lies in at most one lattice face, and only existing faces are listed
    for a in range(1, NS):
        faces = [int(x) for x in pfk[a]]
        assert sum(bin(f).count("1") for f in faces) == bin(faces[0] | faces[1] | faces[2] | faces[3]).count("1")
        assert (faces[0] | faces[1] | faces[2] | faces[3]) < (1 << (4 * npieces[a]))
        for k in range(4):
            for pi in range(6):
                sigs = [(int(fsig[a, k, pi]) >> (8 * i)) & 0xff for i in range(4)]
                mine = [int(pfsig[a, pi, i]) for i in range(12) if (faces[k] >> i) & 1]
                assert [s for s in sigs if s] == mine and all(mine)
    # geometry: the tet across a face sees us across the same face, with the roles of the rank permutations
    # swapped, and exactly one of the two comes later
    for p in range(8):
        for t in range(5):
            for k in range(4):
                g = int(geo[p, t, k])
                assert g & (1 << 18)
                d = [((g >> s) & 3) - 1 for s in (0, 2, 4)]
                t2, k2 = (g >> 6) & 7, (g >> 9) & 3
                assert sum(abs(x) for x in d) <= 1 and (d != [0, 0, 0] or t2 != t)
                p2 = p ^ (1 if d[0] else 0) ^ (2 if d[1] else 0) ^ (4 if d[2] else 0)
                g2 = int(geo[p2, t2, k2])
                assert [((g2 >> s) & 3) - 1 for s in (0, 2, 4)] == [-x for x in d]
                assert ((g2 >> 6) & 7, (g2 >> 9) & 3) == (t, k)
                assert (g >> 11) & 7 == (g2 >> 14) & 7 and (g >> 14) & 7 == (g2 >> 11) & 7
                assert bool(g & (1 << 17)) != bool(g2 & (1 << 17))


def _random_tree(g, rng, depth):
    """A random geom.* expression: primitives under translate / rotate / scale and smooth booleans."""
    if depth == 0 or rng.random() < 0.3:
        k = int(rng.integers(0, 4))
        if k == 0:
            o = g.sphere(float(rng.uniform(0.05, 0.6)))
        elif k == 1:
            o = g.cube(*(float(x) for x in rng.uniform(0.1, 0.9, 3)), float(rng.choice([0.0, 0.02, 0.1])))
        elif k == 2:
            o = g.cylinder(float(rng.uniform(0.2, 1.5)), float(rng.uniform(0.05, 0.4)), float(rng.uniform(0.05, 0.4)),
                           float(rng.choice([0.0, 0.05])))
        else:
            o = g.sphere(float(rng.uniform(0.1, 0.4))).scale(*(float(x) for x in rng.uniform(0.5, 2.0, 3)))
    else:
        a, b = _random_tree(g, rng, depth - 1), _random_tree(g, rng, depth - 1)
        r = float(rng.choice([0.0, 0.03, 0.2]))
        o = [a.union, a.intersection, a.difference][int(rng.integers(0, 3))](b, r)
    for _ in range(int(rng.integers(0, 3))):
        k = int(rng.integers(0, 3))
        v = [float(x) for x in (rng.uniform(-1, 1, 3) if k == 0 else rng.uniform(-3.2, 3.2, 3) if k == 1 else rng.uniform(0.3, 2.5, 3))]
        o = [o.translate, o.rotate, o.scale][k](*v)
    return o


@pytest.mark.parametrize("seed", range(40))
def test_random_scenes_bbox_and_lowering(seed, oracle):
    """Host lowering against the oracle on random scenes: identical bounding boxes (bbox 0.11.2 / implicit3d rules:
    union dilation, cofactor inverse of composed transforms) at every level, and the program compiles within the
    evaluator's limits."""
    rng_a, rng_b = np.random.default_rng(seed), np.random.default_rng(seed)
    ga, gb = mc.Geom(), mc.Geom(oracle.OracleScene())
    a, b = _random_tree(ga, rng_a, 3), _random_tree(gb, rng_b, 3)
    ba, bb = np.array(a.bbox()), np.array(b.bbox())
    assert np.array_equal(ba, bb), (seed, ba, bb)
    if np.all(np.isfinite(ba)) and np.all(ba[3:] > ba[:3]):
        info = ga.backend.program_info(a.node, 0.01)
        assert 0 < info["words"] < (1 << 22) and info["value_depth"] <= 48 and info["point_depth"] <= 6


def _program_words(obj, slack):
    import ctypes as C
    lib = _lib.lib()
    n = lib.mc_tape_program_words(obj.backend.handle, obj.node, float(slack), None, 0)
    assert n > 0
    buf = (C.c_uint64 * n)()
    assert lib.mc_tape_program_words(obj.backend.handle, obj.node, float(slack), buf, n) == n
    return list(buf)


LOWERING_SCRIPTS = {
    "cube": scenes.CUBE_LUA, "sphere": scenes.SPHERE_LUA, "xplicit": scenes.XPLICIT_WITH_TOP_LUA,
    "cyl": "return { c = geom.cylinder(2.0, 0.5, 0.25, 0.1):rotate(0.3, 0.2, 0.1):translate(1, 2, 3):scale(2, 3, 4):volume() }",
    "mix": "a = geom.sphere(0.3):translate(0.2, 0, 0)\nb = geom.cube(0.5, 0.4, 0.3, 0.05):rotate(1, 2, 3)\n"
           "c = geom.plane_hesian(1, 2, 3, 0.1)\nd = geom.plane_3_points(0,0,0.2, 1,0,0.25, 0,1,0.3)\n"
           "return { u = a:union(b, 0.1):intersection(c, 0.05):difference(d, 0.0):volume(), "
           "k = geom.i_cone(0.5):intersection(geom.i_cylinder(0.4), 0.1):scale(1, 2, 1):boundary() }",
}


@pytest.mark.parametrize("name", sorted(LOWERING_SCRIPTS))
def test_lowered_program_evaluates_like_the_oracle(name, oracle):
    """The flat device program the host lowering emits (opcodes, parameters, per-node slack constants, composed
    matrices, jump targets), run by a plain-Python reference interpreter (tests/program_interp.py), gives the
    oracle's recursive approx_value bit for bit — the lowering is checked on the CPU, independently of the CUDA
    evaluator (which tests/test_gpu_parity.py checks against the same oracle on the device)."""
    import program_interp
    oracle.set_libm_mode(1)
    L = oracle.lib()
    a, b = _both(LOWERING_SCRIPTS[name], oracle)
    rng = np.random.default_rng(3)
    for k in sorted(a):
        bb = np.array(a[k].obj.bbox())
        lo = np.where(np.isfinite(bb[:3]), bb[:3], -2.0) - 0.3
        hi = np.where(np.isfinite(bb[3:]), bb[3:], 2.0) + 0.3
        pts = rng.uniform(lo, hi, size=(300, 3))
        pts[:3] = 0.0
        for slack in (0.0, 0.07, 1.0):
            run = program_interp.Interp(_program_words(a[k].obj, slack), L.orc_exp, L.orc_ln)
            want = b[k].obj.backend.eval_points(b[k].obj.node, slack, pts)
            got = np.array([run(float(p[0]), float(p[1]), float(p[2])) for p in pts])
            same = (got.view(np.uint64) == want.view(np.uint64)) | (got == want)
            assert same.all(), (name, k, slack, pts[~same][:3], got[~same][:3], want[~same][:3])


def test_synthetic_scene_program_evaluates_like_the_oracle(oracle):
    import program_interp
    oracle.set_libm_mode(1)
    L = oracle.lib()
    g, go = mc.Geom(), mc.Geom(oracle.OracleScene())
    root, root_o = scenes.synthetic_csg(g, 16, seed=16), scenes.synthetic_csg(go, 16, seed=16)
    g.backend.set_parameters(0.1, 1.0)
    root_o.backend.set_parameters_node(root_o.node, 0.1, 1.0)
    pts = np.random.default_rng(5).uniform(-0.6, 0.6, size=(400, 3))
    slack = 1.0 / 64
    run = program_interp.Interp(_program_words(root, slack), L.orc_exp, L.orc_ln)
    want = root_o.backend.eval_points(root_o.node, slack, pts)
    got = np.array([run(*map(float, p)) for p in pts])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64)) or np.array_equal(got, want)
