"""GPU parity, C5 (spike-and-slab variable selection): CUDA warp-per-replicate kernel vs the oracle, bit-exact."""
import numpy as np
import pytest

from tests import cases
from tests.test_gpu_parity_c1_c2 import assert_same
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R

pytestmark = pytest.mark.gpu
KEYS = ["meetingtime", "iteration", "cost", "uestimator", "mcmcestimator", "correction", "summary"]


@pytest.mark.parametrize("n,p,kappa,s0,lag,k,m", [(50, 8, 0.1, 4, 1, 0, 40), (100, 30, 0.1, 10, 1, 50, 300),
                                                  (80, 70, 1.0, 100, 3, 10, 100), (500, 1000, 2.0, 100, 1, 100, 600)])
def test_c5_estimator_bit_exact(engine, oracle, n, p, kappa, s0, lag, k, m):
    mdl = cases.c5_model(n, p, kappa=kappa, s0=s0, snr=2.0)
    nrep = 48 if p < 1000 else 24
    kw = dict(k=k, m=m, lag=lag, nrep=nrep, max_iterations=5 * m)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=4), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=4), nthreads=8, **kw)
    assert_same(g, c, KEYS)


def test_c5_meetingtimes_full_shape(engine, oracle):
    """BASELINE shape n=500, p=1000, g=p^3, s0=100, kappa=2 (varselection.differentp.R:22-52: meeting times, k=m=0)."""
    mdl = cases.c5_model()
    g = engine.sample_meetingtime(mdl, R.Draws(seed=1), lag=1, nrep=64, max_iterations=30000)
    c = oracle.sample_meetingtime(mdl, R.Draws(seed=1), lag=1, nrep=64, max_iterations=30000, nthreads=8)
    assert_same(g, c, ["meetingtime"])
    assert np.isfinite(g["meetingtime"]).mean() > 0.9


def test_c5_chains_and_replay(engine, oracle):
    mdl = cases.c5_model(60, 12, kappa=0.1, s0=6, snr=2.0)
    kw = dict(m=60, lag=2, stride=900, nrep=10, max_iterations=800)
    g = engine.sample_coupled_chains(mdl, R.Draws(seed=3), **kw)
    c = oracle.sample_coupled_chains(mdl, R.Draws(seed=3), **kw)
    assert_same(g, c, ["meetingtime", "iteration", "cost"])
    for r in range(10):
        nrow = int(c["iteration"][r]) + 1
        assert np.array_equal(g["samples1"][r, :nrow], c["samples1"][r, :nrow])
        assert np.array_equal(g["samples2"][r, :nrow - 2], c["samples2"][r, :nrow - 2])
    # replay tape (only uniforms are consumed by this model)
    rng = np.random.default_rng(2)
    draws = R.Draws(tape_u=rng.uniform(size=(8, 6000)))
    g = engine.sample_unbiasedestimator(mdl, draws, k=5, m=50, lag=1, max_iterations=700, nrep=8)
    c = oracle.sample_unbiasedestimator(mdl, draws, k=5, m=50, lag=1, max_iterations=700, nrep=8)
    assert_same(g, c, KEYS)


@pytest.mark.parametrize("comp,lag,maxit", [(0, 1, 2000), (17, 2, 2000), (29, 1, 120)])
def test_c5_bins_on_the_fly(engine, oracle, comp, lag, maxit):
    """estimator_bin_ (src/estimator_bin.cpp:15-38) of one inclusion indicator, in closed form from the lazy counters;
    the last case censors some replicates before they meet"""
    mdl = cases.c5_model(100, 30, kappa=0.1, s0=10, snr=2.0)
    bins = R.Bins(comp, np.array([-0.5, 0.5, 1.5]))
    kw = dict(k=30, m=150, lag=lag, nrep=48, max_iterations=maxit)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=6), bins=bins, **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=6), bins=bins, nthreads=8, **kw)
    assert_same(g, c, KEYS + ["bins"])
    ok = np.isfinite(g["meetingtime"])
    if maxit >= 150:
        np.testing.assert_allclose(g["bins"][ok].sum(1), 1.0, atol=1e-12)      # the two values carry all the mass
    np.testing.assert_allclose(g["bins"][ok][:, 1], g["uestimator"][ok][:, comp], atol=1e-12)
    # a break inside (0, 1) and values outside all bins
    bins2 = R.Bins(comp, np.array([0.25, 0.75, 2.0]))
    g2 = engine.sample_unbiasedestimator(mdl, R.Draws(seed=6), bins=bins2, **kw)
    c2 = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=6), bins=bins2, nthreads=8, **kw)
    assert_same(g2, c2, KEYS + ["bins"])
