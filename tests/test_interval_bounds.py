"""CPU check of the interval bound the tile / sub-box sign proofs rest on (mc_prune.cuh, BOOL case).

For children with values in [lo_i, hi_i]:
    -r/4 ln sum exp(-lo_i / (r/4))  <=  rvmin(x)  <=  min hi_i
     max lo_i                       <=  rvmax(x)  <=  r/4 ln sum exp(hi_i / (r/4))
for BOTH branches of the reference's rvmin / rvmax (the exact one and the smooth one, with its
`close` switch and its `x < lim` filter).  The functions under test are the oracle's restatement
of implicit3d's Union / Intersection over three axis planes, whose child values are x, y and z.
"""
import numpy as np
import pytest

import moosecad_b200 as mc


def _scene(oracle, kind, r):
    g = mc.Geom(oracle.OracleScene())
    planes = [g.plane_x(0.0), g.plane_y(0.0), g.plane_z(0.0)]
    be = g.backend
    ids = [p.node for p in planes]
    node = be.union(ids, r) if kind == "union" else be.intersection(ids, r)
    be.set_parameters_node(node, 0.1, 1.0)
    return be, node


@pytest.mark.parametrize("kind", ["union", "intersection"])
@pytest.mark.parametrize("r", [0.05, 0.3, 1e-3])
def test_log_sum_exp_bound_holds(oracle, kind, r):
    be, node = _scene(oracle, kind, r)
    rng = np.random.default_rng(7)
    r4 = r / 4.0
    worst = np.inf
    for scale in (4.0 * r, r, 0.2 * r, 0.02 * r):      # boxes from far apart to deep inside the blend
        lo = rng.uniform(-scale, scale, (400, 3))
        hi = lo + rng.uniform(0.0, 0.5 * scale, (400, 3))
        t = rng.uniform(0.0, 1.0, (400, 16, 3))
        t[:, 0] = 0.0                                    # the corners of the box are the tight cases
        t[:, 1] = 1.0
        pts = lo[:, None, :] + t * (hi - lo)[:, None, :]
        val = be.eval_points(node, 1e-3, pts.reshape(-1, 3)).reshape(400, 16)
        if kind == "union":
            m = lo.min(axis=1)
            lower = m - r4 * np.log(np.exp(-(lo - m[:, None]) / r4).sum(axis=1))
            upper = hi.min(axis=1)
            assert (lower <= m + 1e-15).all() and (lower >= m - r4 * np.log(3.0) - 1e-12).all()
        else:
            m = hi.max(axis=1)
            upper = m + r4 * np.log(np.exp((hi - m[:, None]) / r4).sum(axis=1))
            lower = lo.max(axis=1)
            assert (upper >= m - 1e-15).all() and (upper <= m + r4 * np.log(3.0) + 1e-12).all()
        tol = 1e-12 * (1.0 + np.abs(val))
        assert (val >= lower[:, None] - tol).all(), (kind, r, scale)
        assert (val <= upper[:, None] + tol).all(), (kind, r, scale)
        worst = min(worst, float((val - lower[:, None]).min()), float((upper[:, None] - val).min()))
    assert worst > -1e-12


def test_old_bound_was_looser(oracle):
    """The log-sum-exp bound is never below min lo - r/4 ln k, and is tight when one child dominates."""
    r4 = 0.05 / 4.0
    lo = np.array([[0.02, 0.06, 0.4], [0.02, 0.02, 0.02]])
    m = lo.min(axis=1)
    new = m - r4 * np.log(np.exp(-(lo - m[:, None]) / r4).sum(axis=1))
    old = m - r4 * np.log(3.0)
    assert (new >= old - 1e-15).all()
    assert m[0] - new[0] < 0.05 * (m[0] - old[0])      # one dominant child: a fraction of the slack
    assert abs(new[1] - old[1]) < 1e-15                # equal children: the two coincide
