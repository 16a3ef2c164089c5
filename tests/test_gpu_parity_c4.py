"""GPU parity, C4 (Ising parallel tempering): CUDA wavefront sweeps vs the oracle's sequential scan, bit-exact."""
import numpy as np
import pytest

from tests import cases
from tests.test_gpu_parity_c1_c2 import assert_same
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R

pytestmark = pytest.mark.gpu
KEYS = ["meetingtime", "iteration", "cost", "uestimator", "mcmcestimator", "correction", "summary"]


@pytest.mark.parametrize("nchains,pswap,lo,hi", [(1, 0.0, 0.3, 0.3), (4, 0.05, 0.3, 0.4), (16, 0.2, 0.2, 0.35)])
def test_c4_estimator_bit_exact(engine, oracle, nchains, pswap, lo, hi):
    mdl = cases.c4_model(nchains, pswap, lo, hi)
    kw = dict(k=10, m=60, lag=1, nrep=24)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=13), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=13), nthreads=8, **kw)
    assert_same(g, c, KEYS)


def test_c4_lag_and_censoring(engine, oracle):
    mdl = cases.c4_model(3, 0.1, 0.3, 0.45)
    kw = dict(k=4, m=30, lag=3, max_iterations=40, nrep=40)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=5), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=5), nthreads=8, **kw)
    assert_same(g, c, KEYS)
    assert np.isinf(g["meetingtime"]).any()


def test_c4_meetingtimes_full_ladder(engine, oracle):
    """BASELINE shape: 16 temperatures 0.30..0.55, pswap 0.01 (run.batch.ising.R:169-171); capped for test time."""
    mdl = cases.c4_model()
    g = engine.sample_meetingtime(mdl, R.Draws(seed=1), lag=1, max_iterations=300, nrep=12)
    c = oracle.sample_meetingtime(mdl, R.Draws(seed=1), lag=1, max_iterations=300, nrep=12, nthreads=8)
    assert_same(g, c, ["meetingtime"])


def test_c4_chains_of_natural_statistics(engine, oracle):
    mdl = cases.c4_model(4, 0.1, 0.3, 0.4)
    kw = dict(m=50, lag=2, stride=600, nrep=10, max_iterations=500)
    g = engine.sample_coupled_chains(mdl, R.Draws(seed=2), **kw)
    c = oracle.sample_coupled_chains(mdl, R.Draws(seed=2), **kw)
    assert_same(g, c, ["meetingtime", "iteration", "cost"])
    for r in range(10):
        n = int(c["iteration"][r]) + 1
        assert np.array_equal(g["samples1"][r, :n], c["samples1"][r, :n])
        assert np.array_equal(g["samples2"][r, :n - 2], c["samples2"][r, :n - 2])
    assert np.array_equal(engine.h_bar(g, A.UBM_H_IDENTITY, k=5, m=50), oracle.h_bar(c, A.UBM_H_IDENTITY, k=5, m=50))


def test_c4_replay_tape(engine, oracle):
    mdl = cases.c4_model(2, 0.1, 0.3, 0.35)
    nrep, steps = 4, 40
    rng = np.random.default_rng(7)
    draws = R.Draws(tape_u=rng.uniform(size=(nrep, 2 * 2 * 1024 + (steps + 2) * (1 + 2 * 1024))))
    g = engine.sample_unbiasedestimator(mdl, draws, k=2, m=10, lag=1, max_iterations=steps, nrep=nrep)
    c = oracle.sample_unbiasedestimator(mdl, draws, k=2, m=10, lag=1, max_iterations=steps, nrep=nrep)
    assert_same(g, c, KEYS)
