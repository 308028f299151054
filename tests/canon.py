"""Canonical comparison form of a tet mesh.

Neither the reference nor the GPU path has a reproducible vertex / element order (HashMap and
HashSet iteration in the reference, scan order on the device), so meshes are compared after
relabelling: vertices sorted by their coordinates, every tet rotated by the EVEN permutation
that makes it lexicographically smallest (orientation preserved), tets sorted, side-sets as
sorted lists of sorted node triples.
"""
import itertools

import numpy as np

FACE = np.array([[0, 1, 2], [0, 2, 3], [0, 3, 1], [1, 3, 2]])  # Tet4::face, mesh/src/lib.rs:662


def _even_perms():
    out = []
    for p in itertools.permutations(range(4)):
        inv = sum(1 for i in range(4) for j in range(i + 1, 4) if p[i] > p[j])
        if inv % 2 == 0:
            out.append(p)
    return np.array(out)


EVEN = _even_perms()


def canonical(coords, tets, sidesets=None):
    coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
    tets = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    order = np.lexsort((coords[:, 2], coords[:, 1], coords[:, 0]))
    rank = np.empty(len(coords), dtype=np.int64)
    rank[order] = np.arange(len(coords))
    c = coords[order]
    t = rank[tets] if len(tets) else tets
    if len(t):
        cand = t[:, EVEN]                                  # (m, 12, 4)
        # smallest label first (3 even permutations do that), then the smallest second label
        first = cand[..., 0]
        ok = first == first.min(axis=1, keepdims=True)
        second = np.where(ok, cand[..., 1], np.iinfo(np.int64).max)
        best = np.argmin(second, axis=1)
        ct = cand[np.arange(len(t)), best]
        ct = ct[np.lexsort((ct[:, 3], ct[:, 2], ct[:, 1], ct[:, 0]))]
    else:
        ct = t
    out_sets = {}
    for name, es in (sidesets or {}).items():
        es = np.asarray(es, dtype=np.int64).reshape(-1, 2)
        if len(es):
            tri = np.sort(t[es[:, 0][:, None], FACE[es[:, 1]]], axis=1)
            tri = tri[np.lexsort((tri[:, 2], tri[:, 1], tri[:, 0]))]
        else:
            tri = np.zeros((0, 3), dtype=np.int64)
        out_sets[name] = tri
    return c, ct, out_sets


def assert_same_mesh(a, b, bitwise=True):
    """a, b: (coords, tets, sidesets) triples in canonical form."""
    ca, ta, sa = a
    cb, tb, sb = b
    assert ca.shape == cb.shape, f"vertex count {ca.shape} != {cb.shape}"
    assert ta.shape == tb.shape, f"tet count {ta.shape} != {tb.shape}"
    if bitwise:
        # -0.0 and +0.0 are the same position
        assert np.array_equal(ca, cb), f"coords differ at {np.argwhere(ca != cb)[:5]}"
    else:
        np.testing.assert_allclose(ca, cb, rtol=1e-12, atol=1e-12)
    assert np.array_equal(ta, tb), f"tets differ at rows {np.argwhere((ta != tb).any(axis=1))[:5].ravel()}"
    assert set(sa) == set(sb)
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), f"side-set {k} differs"
