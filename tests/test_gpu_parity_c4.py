"""GPU parity, C4 (Ising parallel tempering): the CUDA row-at-a-time sweeps (carry-chain scan, one lane per lattice
pair) vs the oracle's site-by-site sequential scan, bit-exact."""
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


@pytest.mark.parametrize("nchains,nrep", [(12, 7), (32, 3), (5, 13), (2, 33)])
def test_c4_ragged_groups(engine, oracle, nchains, nrep):
    """Ladders that do not divide the warp (idle lanes), a full-warp ladder, replicate counts that leave the last group
    of a warp part empty."""
    mdl = cases.c4_model(nchains, 0.15, 0.25, 0.4)
    kw = dict(k=3, m=25, lag=2, nrep=nrep, max_iterations=60)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=21), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=21), nthreads=8, **kw)
    assert_same(g, c, KEYS)


def test_c4_negative_and_extreme_betas(engine, oracle):
    """Antiferromagnetic temperatures (acceptance thresholds decrease with the neighbour count: the serial recurrence
    instead of the carry chain) and inverse temperatures so large that probabilities round to 0 / 1 in 32 bits."""
    for betas in ([-0.3, -0.1, 0.2], [0.0, 3.0, 6.0]):
        mdl = R.IsingPT(np.array(betas), 0.1)
        kw = dict(k=2, m=20, lag=1, nrep=9, max_iterations=40)
        g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=3), **kw)
        c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=3), nthreads=8, **kw)
        assert_same(g, c, KEYS)


def test_c4_bins_on_the_fly(engine, oracle):
    """histogram of one temperature's natural statistic (estimator_bin_, src/estimator_bin.cpp:15-38) inside the kernel"""
    mdl = cases.c4_model(4, 0.1, 0.3, 0.4)
    bins = R.Bins(2, np.linspace(-100.0, 1500.0, 33))
    kw = dict(k=5, m=40, lag=1, nrep=20, max_iterations=300)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=8), bins=bins, **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=8), bins=bins, nthreads=8, **kw)
    assert_same(g, c, KEYS + ["bins"])
    assert g["bins"].sum() > 0


def test_c4_replay_tape_two_replicates_per_warp(engine, oracle):
    """tape mode with several replicates per warp, tapes of exactly the needed length for some replicates only"""
    mdl = cases.c4_model(3, 0.2, 0.3, 0.35)
    nrep, steps = 11, 12
    rng = np.random.default_rng(9)
    draws = R.Draws(tape_u=rng.uniform(size=(nrep, 2 * 3 * 1024 + steps * (1 + 3 * 1024))))
    g = engine.sample_unbiasedestimator(mdl, draws, k=2, m=8, lag=1, max_iterations=steps, nrep=nrep)
    c = oracle.sample_unbiasedestimator(mdl, draws, k=2, m=8, lag=1, max_iterations=steps, nrep=nrep)
    assert_same(g, c, KEYS)


def test_c4_swap_counters(engine, oracle):
    """nswap_attempts / nswap_accepts1 / nswap_accepts2 of the PT kernels (R/ising.R:48-49,76,89-91,137-140), summed over a run"""
    mdl = cases.c4_model(6, 0.3, 0.25, 0.45)
    kw = dict(k=5, m=40, lag=2, nrep=37, max_iterations=200)
    oracle.ising_swap_counts(6, reset=True)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=31), nthreads=8, **kw)
    want = oracle.ising_swap_counts(6, reset=True)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=31), **kw)
    got = engine.ising_swap_counts(6)
    assert_same(g, c, KEYS)
    assert got["nswap_attempts"] == want["nswap_attempts"] > 0
    assert np.array_equal(got["nswap_accepts1"], want["nswap_accepts1"]) and got["nswap_accepts1"].sum() > 0
    assert np.array_equal(got["nswap_accepts2"], want["nswap_accepts2"]) and got["nswap_accepts2"].sum() > 0
