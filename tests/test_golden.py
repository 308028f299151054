"""Golden fixtures (tests/golden/*.npz, written by scripts/make_golden.py).

The Rust reference cannot be run here, so the vectors come from the ORACLE after it had been pinned to the
reference's own known answers; they freeze that state.  CPU: the oracle still reproduces them bit for bit.
GPU: the device path (through the C ABI) reproduces them without the oracle in the loop — canonical mesh,
Mesh::boundary, side-sets, statistics, and 512 SDF values per scene, all bit-exact."""
import os
import sys

import numpy as np
import pytest

import canon

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import make_golden  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def test_every_case_has_a_fixture():
    assert sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz")) == sorted(make_golden.CASES)


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden(name, oracle):
    kind, n, err = make_golden.CASES[name]
    want, got = _load(name), make_golden.oracle_case(kind, n, err)
    oracle.set_libm_mode(1)
    assert sorted(want) == sorted(got)
    for k in want:
        a, b = np.asarray(want[k]), np.asarray(got[k])
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if a.dtype == np.float64:
            assert np.array_equal(a.reshape(-1).view(np.uint64), b.reshape(-1).view(np.uint64)) or np.array_equal(a, b), k
        else:
            assert np.array_equal(a, b), k


def test_config1_golden_is_the_hand_traced_answer():
    # examples/cube.lua at --tessellation-resolution 1.0 (BASELINE config 1): 8 vertices, 5 tets, 12 surface sides,
    # top / bottom = 2 + 2 sides, nothing trimmed
    g = _load("cube_res1")
    assert g["coords"].shape == (8, 3) and g["tets"].shape == (5, 4) and g["set___boundary__"].shape == (12, 3)
    assert g["set_top"].shape == (2, 3) and g["set_bottom"].shape == (2, 3)
    assert g["dim"].tolist() == [1, 1, 1] and g["stats"].tolist() == [0] * 7
    assert sorted(map(tuple, g["coords"].tolist())) == sorted((x, y, z) for x in (-.5, .5) for y in (-.5, .5) for z in (-.5, .5))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_device_reproduces_golden(name, ctx):
    import moosecad_b200 as mc
    kind, n, err = make_golden.CASES[name]
    g = _load(name)
    vol, bnd, _ = make_golden.build(kind, None)
    vol.backend.set_parameters(0.1, 1.0)
    res = float(g["resolution"])
    mesh = mc.IsosurfaceStuffing(vol, res, err, ctx=ctx, allow_diverged=True).tessellate()
    sets = {"__boundary__": sorted(mesh.boundary().sides())}
    for sname, b in bnd.items():
        sets[sname] = sorted(mesh.add_boundary_side_set(sname, b).sides())
    c, t, s = canon.canonical(mesh.vertices(), mesh.elems(), sets)
    want = (g["coords"], g["tets"], {k[4:]: g[k] for k in g if k.startswith("set_")})
    canon.assert_same_mesh((c, t, s), want, bitwise=True)
    st = mesh.stats()
    assert st["dim"] == g["dim"].tolist() and st["stats"] == g["stats"].tolist()
    assert st["ntets_stage"] == g["ntets_stage"].tolist()
    assert [st["n_warped"], st["n_cut"]] == g["n_warped_cut"].tolist()
    got = mc.ObjectAdaptor(vol, res, ctx).value(g["points"])
    assert np.array_equal(got.view(np.uint64), g["values"].view(np.uint64)) or np.array_equal(got, g["values"])
