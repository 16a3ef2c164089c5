"""GPU, BASELINE.json's full sizes for C3, C4, C5 (C2: test_gpu_parity_c1_c2.py::test_c2_full_size_properties), where the
oracle is too slow to be the checker: size-independent properties — the cost / iteration identities of
R/coupled_chains.R:82, uestimator = (mcmc + correction) / (m - k + 1), range and ordering facts of each model, and
independence of every per-replicate result from how the replicates are split over launches (work stealing, rep_offset)."""
import numpy as np
import pytest

from tests import cases
from unbiasedmcmc_b200 import registry as R

pytestmark = pytest.mark.gpu


def _common(engine, mdl, k, m, nrep, seed, split):
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=seed), k=k, m=m, lag=1, nrep=nrep)
    tau, it, cost = g["meetingtime"], g["iteration"], g["cost"]
    assert np.isfinite(tau).all() and tau.min() >= 2
    assert np.array_equal(it, np.maximum(tau, m).astype(np.int64))
    assert np.array_equal(cost, 1 + 2 * (tau - 1) + np.maximum(0, it - tau))
    np.testing.assert_allclose(g["uestimator"], g["mcmcestimator"] + g["correction"], rtol=1e-13, atol=1e-13)
    assert g["summary"][0] == nrep and g["summary"][1] == 0
    a = engine.sample_unbiasedestimator(mdl, R.Draws(seed=seed), k=k, m=m, lag=1, nrep=split)
    b = engine.sample_unbiasedestimator(mdl, R.Draws(seed=seed), k=k, m=m, lag=1, nrep=nrep - split, rep_offset=split)
    for key in ("uestimator", "meetingtime", "cost"):
        assert np.array_equal(g[key], np.concatenate([a[key], b[key]])), key
    return g


def test_c3_full_size_properties(engine):
    """germancredit shape: n = 1000, p = 49, k = 110, m = 1100."""
    g = _common(engine, cases.c3_model(), 110, 1100, 74, seed=5, split=20)
    u = g["uestimator"]
    assert np.isfinite(u).all()
    # meeting times on this synthetic design exceed k = 110 (mean ~ 540), so single estimators are noisy (sd 2-8 per
    # coefficient); their average still sits at the scale of the posterior means
    assert np.abs(u.mean(0)).max() < 10.0


def test_c4_full_size_properties(engine):
    """32 x 32 Ising, 16 temperatures in [0.3, 0.55], k = 1000, m = 2000: the natural statistic (sum over the 2048 edges)
    lies in [-2048, 2048] for every MCMC average, which increases with the inverse temperature."""
    g = _common(engine, cases.c4_model(), 1000, 2000, 74, seed=6, split=30)
    mc = g["mcmcestimator"]
    assert mc.min() >= -2048 and mc.max() <= 2048
    # the MCMC part (the corrections are large and noisy here: meeting times are in the thousands, k = 1000)
    assert np.all(np.diff(mc.mean(0)[::3]) > 0)      # every third temperature: increments well above the noise
    # integer arithmetic end to end: (m - k + 1) * mcmcestimator is a sum of integers
    t = mc * 1001
    assert np.max(np.abs(t - np.round(t))) < 1e-6


def test_c5_full_size_properties(engine):
    """n = 500, p = 1000, s0 = 100, k = 7500, m = 15000: inclusion frequencies lie in [0, 1], the model size stays
    within s0, and the exact-integer counters make (m - k + 1) * mcmcestimator integral."""
    g = _common(engine, cases.c5_model(), 7500, 15000, 148, seed=7, split=50)
    mc = g["mcmcestimator"]
    assert mc.min() >= 0.0 and mc.max() <= 1.0
    assert mc.sum(1).max() <= 100.0
    t = mc * 7501
    assert np.max(np.abs(t - np.round(t))) < 1e-6
    incl = g["uestimator"].mean(0)
    assert incl[:10].mean() > 5 * max(incl[10:].mean(), 1e-3)        # the ten true signals stand out
