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
