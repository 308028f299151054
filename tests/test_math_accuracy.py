"""The portable exp / ln (oracle/oracle_math.h, mirrored on the device in
moosecad_b200/csrc/mc_math.cuh) against mpmath: < 1 ulp.  CPU only."""
import numpy as np
import pytest

mpmath = pytest.importorskip("mpmath")


def _ulp_err(got, want_mp):
    want = float(want_mp)
    if want == 0.0 or not np.isfinite(want):
        return 0.0 if got == want else np.inf
    return abs(mpmath.mpf(got) - want_mp) / mpmath.mpf(float(np.spacing(abs(want))))


def test_exp_ln_below_one_ulp(oracle):
    mpmath.mp.prec = 200
    L = oracle.lib()
    oracle.set_libm_mode(1)
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-60, 20, 4000), rng.uniform(-1, 1, 2000), rng.uniform(-1e-3, 1e-3, 500),
                         [0.0, -0.0, 0.34657359027997264, -0.34657359027997264, 709.0, -745.0, 1e-300]])
    worst = max(_ulp_err(L.orc_exp(float(x)), mpmath.exp(mpmath.mpf(float(x)))) for x in xs)
    assert worst < 1.0, worst
    ys = np.concatenate([np.exp(rng.uniform(-80, 80, 4000)), 1.0 + rng.uniform(-0.3, 0.4, 2000),
                         1.0 + rng.uniform(-1e-6, 1e-6, 1000), [1.0, 0.5, 2.0, 5e-324, 1e308]])
    worst = max(_ulp_err(L.orc_ln(float(y)), mpmath.log(mpmath.mpf(float(y)))) for y in ys)
    assert worst < 1.0, worst
    assert L.orc_exp(float("inf")) == float("inf") and L.orc_exp(float("-inf")) == 0.0
    assert L.orc_ln(0.0) == float("-inf") and np.isnan(L.orc_ln(-1.0))
