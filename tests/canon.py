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


# ---- numpy mirror of mc_mesh_checksum (moosecad_b200/csrc/mc_hash.cuh) ----------------------
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(30)
    x = x * np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x = x * np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def vertex_hashes(coords):
    c = np.asarray(coords, dtype=np.float64).reshape(-1, 3) + 0.0   # -0.0 -> +0.0
    b = c.view(np.uint64)
    h = np.full(len(c), 0x9E3779B97F4A7C15, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k in range(3):
            h = _mix64(h + b[:, k])
    return h


def mesh_checksum(coords, tets):
    """(vertex checksum, tet checksum): what mc_mesh_checksum returns in out[0], out[1]."""
    hv = vertex_hashes(coords)
    tets = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    with np.errstate(over="ignore"):
        vsum = int(np.add.reduce(hv, dtype=np.uint64)) if len(hv) else 0
        if not len(tets):
            return vsum, 0
        cand = hv[tets][:, EVEN]                      # (m, 12, 4) hash sequences
        # lexicographic minimum over the 12 even rotations
        best = cand[:, 0, :].copy()
        for p in range(1, 12):
            s = cand[:, p, :]
            ne = s != best
            first = np.argmax(ne, axis=1)
            rows = np.arange(len(s))
            less = ne.any(axis=1) & (s[rows, first] < best[rows, first])
            best[less] = s[less]
        t = np.full(len(tets), 0xD1B54A32D192ED03, dtype=np.uint64)
        for k in range(4):
            t = _mix64(t + best[:, k])
        return vsum, int(np.add.reduce(t, dtype=np.uint64))


def mesh_diff(coords_a, tets_a, coords_b, tets_b, decimals=9):
    """Topology difference of two meshes whose vertex positions may differ in the last bits:
    vertices are matched after rounding to `decimals`, tets are compared as oriented node sets
    (canonical even rotation).  Returns a dict of counts."""
    ca = np.round(np.asarray(coords_a, dtype=np.float64).reshape(-1, 3), decimals) + 0.0
    cb = np.round(np.asarray(coords_b, dtype=np.float64).reshape(-1, 3), decimals) + 0.0
    allc, inv = np.unique(np.concatenate([ca, cb]), axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    ia, ib = inv[:len(ca)], inv[len(ca):]

    def canon_tets(t, ids):
        t = ids[np.asarray(t, dtype=np.int64).reshape(-1, 4)]
        if not len(t):
            return t
        cand = t[:, EVEN]
        first = cand[..., 0]
        ok = first == first.min(axis=1, keepdims=True)
        second = np.where(ok, cand[..., 1], np.iinfo(np.int64).max)
        return cand[np.arange(len(t)), np.argmin(second, axis=1)]

    ta, tb = canon_tets(tets_a, ia), canon_tets(tets_b, ib)
    va, vb = np.unique(ia), np.unique(ib)
    rows = np.concatenate([ta, tb])
    tag = np.concatenate([np.zeros(len(ta), dtype=np.int64), np.ones(len(tb), dtype=np.int64)])
    if len(rows):
        _, idx, cnt = np.unique(rows, axis=0, return_index=True, return_counts=True)
        only = idx[cnt == 1]
        t_only_a, t_only_b = int((tag[only] == 0).sum()), int((tag[only] == 1).sum())
    else:
        t_only_a = t_only_b = 0
    return {"vertices_a": len(ca), "vertices_b": len(cb),
            "vertices_only_a": int(np.setdiff1d(va, vb).size), "vertices_only_b": int(np.setdiff1d(vb, va).size),
            "tets_a": len(ta), "tets_b": len(tb), "tets_only_a": t_only_a, "tets_only_b": t_only_b}
