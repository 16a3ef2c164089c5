"""CPU: the oracle pinned by the identities and anchors the reference's own check scripts establish
(SURVEY.md §4 / §8c).  The reference ships no golden vectors and cannot run here, so these are the pins."""
import numpy as np
from scipy import stats

from tests import cases
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R


def test_weights_identity_hbar_equals_bruteforce_average(oracle):
    """inst/check/check_hbar.R:89-178: H_bar == average of H_l for l = k..m (brute force), lag 1."""
    mdl = cases.c1_model()
    k, m = 5, 40
    ch = oracle.sample_coupled_chains(mdl, R.Draws(seed=5), m=m, lag=1, stride=4096, nrep=50)
    hb = oracle.h_bar(ch, A.UBM_H_IDENTITY, k=k, m=m)[:, 0]
    for r in range(50):
        s1, s2 = ch["samples1"][r, :, 0], ch["samples2"][r, :, 0]
        tau = int(ch["meetingtime"][r])
        hl = []
        for l in range(k, m + 1):
            v = s1[l]
            for t in range(l + 1, tau):
                v += s1[t] - s2[t - 1]
            hl.append(v)
        assert abs(np.mean(hl) - hb[r]) < 1e-10


def test_on_the_fly_equals_hbar_of_stored_chains(oracle):
    """inst/check/check_lags.R:71-97 checks this statistically; under shared draws it is an identity."""
    mdl = cases.c1_model()
    bins = cases.c1_bins()
    for lag, k, m in [(1, 0, 30), (10, 10, 50), (50, 0, 60), (7, 20, 20)]:
        ch = oracle.sample_coupled_chains(mdl, R.Draws(seed=9), m=m, lag=lag, stride=8192, nrep=64)
        e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=9), h_kind=A.UBM_H_BOTH, bins=bins, k=k, m=m, lag=lag, nrep=64)
        assert np.array_equal(ch["meetingtime"], e["meetingtime"])
        assert np.array_equal(ch["cost"], e["cost"]) and np.array_equal(ch["iteration"], e["iteration"])
        np.testing.assert_allclose(oracle.h_bar(ch, A.UBM_H_BOTH, k=k, m=m), e["uestimator"], rtol=1e-12, atol=1e-12)
        assert np.array_equal(oracle.estimator_bins(ch, bins, k=k, m=m), e["bins"])
        np.testing.assert_allclose(e["mcmcestimator"] + e["correction"], e["uestimator"], rtol=1e-12, atol=1e-13)


def test_cost_and_time_indexing(oracle):
    """R/coupled_chains.R:82: cost = lag + 2 (tau - lag) + max(0, iteration - tau); tau >= lag + 1."""
    mdl = cases.c1_model()
    lag, m = 7, 30
    e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=3), k=3, m=m, lag=lag, nrep=500)
    tau, it, cost = e["meetingtime"], e["iteration"], e["cost"]
    assert np.all(tau >= lag + 1)
    assert np.array_equal(it, np.maximum(tau, m).astype(np.int64))
    assert np.array_equal(cost, lag + 2 * (tau - lag) + np.maximum(0, it - tau))


def test_censoring_semantics(oracle):
    """R/coupled_chains.R:57-58: max_iterations stops the loop, meetingtime (and cost) stay Inf."""
    mdl = cases.c1_model()
    e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=3), k=2, m=10, lag=2, max_iterations=12, nrep=800)
    cens = np.isinf(e["meetingtime"])
    assert cens.any() and (~cens).any()
    assert np.all(e["iteration"][cens] == 12) and np.all(np.isinf(e["cost"][cens]))
    assert e["summary"][0] == (~cens).sum() and e["summary"][1] == cens.sum()


def test_c1_readme_anchors(oracle):
    """README.md:146,171: mean cost 2072.788, 95% CI 0.0055 +/- 0.1459 from 500 replicates (sd ~ 1.664)."""
    mdl = cases.c1_model()
    e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=1), bins=cases.c1_bins(), k=500, m=2000, lag=500, nrep=6000,
                                        nthreads=8)
    cost, h = e["cost"], e["uestimator"][:, 0]
    assert abs(cost.mean() - 2072.788) < 4 * cost.std() / np.sqrt(500) + 4 * cost.std() / np.sqrt(6000)
    assert abs(h.mean()) < 4 * h.std() / np.sqrt(6000)                 # target mean is 0
    assert 1.45 < h.std() < 1.85                                         # README: 1.664 from 500 replicates
    # histogram estimates a density that integrates to 1 and matches the mixture
    prop = e["bins"].mean(0)
    assert abs(prop.sum() - 1.0) < 0.01
    mids = (cases.c1_bins().breaks[:-1] + cases.c1_bins().breaks[1:]) / 2
    dens = 0.5 * stats.norm.pdf(mids, -4, 1) + 0.5 * stats.norm.pdf(mids, 4, 1)
    se = e["bins"].std(0) / np.sqrt(6000)
    assert np.all(np.abs(prop / 0.2 - dens) < 5 * se / 0.2 + 2e-3)


def test_c1_indicator_estimand(oracle):
    """inst/reproducebimodal/bimodal.mcmc.run.R:36: E[1{x > 3}] = 0.5 * Phi(1) = 0.4207 — via the histogram bins."""
    mdl = cases.c1_model()
    bins = R.Bins(0, np.array([3.0, 1e9]))
    e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=2), bins=bins, k=500, m=2000, lag=500, nrep=3000, nthreads=8)
    b = e["bins"][:, 0]
    assert abs(b.mean() - 0.5 * stats.norm.cdf(1.0)) < 4 * b.std() / np.sqrt(3000)


def test_normal_target_moments(oracle):
    """vignettes/introduction-normaltarget.Rmd:242,271: N(0,1) target => E x = 0, E x^2 = 1 (lag 10, k 10, m 50 as
    in inst/check/check_lags.R)."""
    mdl = R.NormalMixtureRWMH([1.0], [0.0], [1.0], 1.0, 5.0, 2.0)
    e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=4), h_kind=A.UBM_H_BOTH, k=10, m=50, lag=10, nrep=20000, nthreads=8)
    u = e["uestimator"]
    assert abs(u[:, 0].mean()) < 4 * u[:, 0].std() / np.sqrt(20000)
    assert abs(u[:, 1].mean() - 1.0) < 4 * u[:, 1].std() / np.sqrt(20000)


def test_mvn_target_moments_and_affine_invariance(oracle):
    d = 6
    mdl = cases.c2_model(d, "dense", "offset")
    e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=5), h_kind=A.UBM_H_BOTH, k=60, m=300, lag=1, nrep=4000, nthreads=8)
    u = e["uestimator"]
    Sigma = cases.mvn_sigma(d, "dense")
    se = u.std(0) / np.sqrt(4000)
    assert np.all(np.abs(u[:, :d].mean(0)) < 4.5 * se[:d])
    assert np.all(np.abs(u[:, d:].mean(0) - np.diag(Sigma)) < 4.5 * se[d:])
    # RW-MH with proposal Sigma/d started from the target is affine invariant: dense and AR(0.5) targets give the same
    # meeting times for the same draws (up to rounding in rare near-ties)
    a = oracle.sample_meetingtime(cases.c2_model(20, "dense"), R.Draws(seed=6), nrep=200)["meetingtime"]
    b = oracle.sample_meetingtime(cases.c2_model(20, "sparse"), R.Draws(seed=6), nrep=200)["meetingtime"]
    assert (a == b).mean() > 0.97


def test_replay_tapes_reproduce_philox_run(oracle):
    mdl = cases.c1_model()
    nrep, L = 32, 3000
    tu = np.stack([oracle.stream(77, r, 0, L) for r in range(nrep)])
    tn = np.stack([oracle.stream(77, r, 1, L) for r in range(nrep)])
    kw = dict(bins=cases.c1_bins(), k=100, m=400, lag=50, nrep=nrep)
    a = oracle.sample_unbiasedestimator(mdl, R.Draws(tape_u=tu, tape_n=tn), **kw)
    b = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=77), **kw)
    for key in ("uestimator", "meetingtime", "cost", "bins"):
        assert np.array_equal(a[key], b[key])


def test_tape_exhaustion_reported(oracle):
    import pytest
    mdl = cases.c1_model()
    rng = np.random.default_rng(1)
    with pytest.raises(oracle.OracleError) as e:
        oracle.sample_unbiasedestimator(mdl, R.Draws(tape_u=rng.uniform(size=(4, 10)), tape_n=rng.standard_normal((4, 10))),
                                        k=10, m=100, lag=1, nrep=4)
    assert e.value.status == A.UBM_ERR_TAPE


def test_argument_guards(oracle):
    import pytest
    mdl = cases.c1_model()
    for kw in (dict(k=5, m=3, lag=1), dict(k=0, m=3, lag=4), dict(k=0, m=3, lag=0)):
        with pytest.raises(oracle.OracleError) as e:
            oracle.sample_unbiasedestimator(mdl, R.Draws(seed=1), nrep=2, **kw)
        assert e.value.status == A.UBM_ERR_INVALID


def test_signed_measure_reproduces_hbar(oracle):
    """c_chains_to_measure_as_list_ (src/c_chains_to_measure.cpp:38-84): number of atoms (m-k+1) + 2 max(0, tau-(k+lag)),
    MCMC weights 1/(m-k+1) summing to one, correction weights in +/- pairs summing to zero, and
    sum_n w_n h(Z_n) == H_bar(c_chains, h, k, m) for any h (inst/check/check_hbar.R compares the two the same way)."""
    mdl = cases.c1_model()
    for lag, k, m in ((1, 3, 40), (5, 10, 60), (7, 0, 25)):
        ch = oracle.sample_coupled_chains(mdl, R.Draws(seed=17), m=m, lag=lag, stride=4000, nrep=40)
        ms = oracle.c_chains_to_measure(ch, k=k, m=m)
        hb = oracle.h_bar(ch, A.UBM_H_BOTH, k=k, m=m)
        for r, M in enumerate(ms):
            tau = int(ch["meetingtime"][r])
            assert len(M["weights"]) == (m - k + 1) + 2 * max(0, tau - (k + lag))
            assert M["MCMC"].sum() == m - k + 1 and np.allclose(M["weights"][M["MCMC"]], 1.0 / (m - k + 1))
            cw = M["weights"][~M["MCMC"]]
            assert np.array_equal(cw[0::2], -cw[1::2]) and np.all(cw[0::2] > 0)
            x = M["atoms"][:, 0]
            np.testing.assert_allclose([np.dot(M["weights"], x), np.dot(M["weights"], x * x)], hb[r], rtol=1e-11, atol=1e-12)
