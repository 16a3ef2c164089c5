"""GPU parity, Bayesian lasso (R/bayesianlasso.R, src/blassoutil.cpp): the CUDA kernel vs the oracle, bit-exact."""
import numpy as np
import pytest

from tests import cases
from tests.test_gpu_parity_c1_c2 import assert_same
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R

pytestmark = pytest.mark.gpu
KEYS = ["meetingtime", "iteration", "cost", "uestimator", "mcmcestimator", "correction", "summary"]


@pytest.mark.parametrize("n,p,lam,lag,k,m,h", [(60, 5, 1.0, 1, 0, 30, A.UBM_H_IDENTITY), (120, 8, 1.0, 1, 10, 80, A.UBM_H_BOTH),
                                               (200, 33, 0.5, 3, 5, 40, A.UBM_H_IDENTITY), (442, 60, 2.0, 1, 5, 30, A.UBM_H_SQUARE)])
def test_blasso_estimator_bit_exact(engine, oracle, n, p, lam, lag, k, m, h):
    mdl = cases.blasso_model(n, p, lam)
    kw = dict(h_kind=h, k=k, m=m, lag=lag, nrep=20 if p < 40 else 8, max_iterations=5000)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=4), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=4), nthreads=8, **kw)
    assert_same(g, c, KEYS)
    assert np.isfinite(g["meetingtime"]).all()


def test_blasso_meetingtimes_chains_and_bins(engine, oracle):
    mdl = cases.blasso_model(120, 8, 1.0)
    g = engine.sample_meetingtime(mdl, R.Draws(seed=1), lag=2, nrep=40, max_iterations=2000)
    c = oracle.sample_meetingtime(mdl, R.Draws(seed=1), lag=2, nrep=40, max_iterations=2000, nthreads=8)
    assert_same(g, c, ["meetingtime"])
    kw = dict(m=40, lag=2, stride=2100, nrep=6, max_iterations=2000)
    cg = engine.sample_coupled_chains(mdl, R.Draws(seed=2), **kw)
    cc = oracle.sample_coupled_chains(mdl, R.Draws(seed=2), **kw)
    assert_same(cg, cc, ["meetingtime", "iteration", "cost"])
    for r in range(6):
        nn = int(cc["iteration"][r]) + 1
        assert np.array_equal(cg["samples1"][r, :nn], cc["samples1"][r, :nn])
        assert np.array_equal(cg["samples2"][r, :nn - 2], cc["samples2"][r, :nn - 2])
    assert np.array_equal(engine.h_bar(cg, A.UBM_H_IDENTITY, k=5, m=40), oracle.h_bar(cc, A.UBM_H_IDENTITY, k=5, m=40))
    bins = R.Bins(16, np.linspace(0.0, 12.0, 25))          # sigma2 = component 2 p
    kw = dict(k=5, m=60, lag=1, nrep=24, max_iterations=2000)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=8), bins=bins, **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=8), bins=bins, nthreads=8, **kw)
    assert_same(g, c, KEYS + ["bins"])
    assert g["bins"].sum() > 0


def test_blasso_replay_tapes(engine, oracle):
    mdl = cases.blasso_model(60, 5, 1.0)
    rng = np.random.default_rng(3)
    nrep = 6
    draws = R.Draws(tape_u=rng.uniform(size=(nrep, 4000)), tape_n=rng.standard_normal((nrep, 4000)))
    kw = dict(k=2, m=12, lag=1, nrep=nrep, max_iterations=40)
    g = engine.sample_unbiasedestimator(mdl, draws, **kw)
    c = oracle.sample_unbiasedestimator(mdl, draws, **kw)
    assert_same(g, c, KEYS)
    from unbiasedmcmc_b200.engine import UbmError
    short = R.Draws(tape_u=rng.uniform(size=(nrep, 30)), tape_n=rng.standard_normal((nrep, 30)))
    with pytest.raises(UbmError):
        engine.sample_unbiasedestimator(mdl, short, **kw)
