"""CPU: oracle pins for C3 (Polya-Gamma logistic) and C4 (Ising PT) — analytic moments and MCMC cross-checks
(PG(1,z) mean tanh(z/2)/(2z) and variance from src/PolyaGamma.cpp:232-263; unbiased estimator vs a
long single chain as in inst/check/check_ising.R:112-164)."""
import numpy as np
from scipy.special import log_ndtr

from tests import cases
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R


def test_pnorm_log_matches_scipy(oracle):
    x = np.concatenate([np.linspace(-38, 9, 100001), np.random.default_rng(0).uniform(-1, 1, 50000)])
    y, ref = oracle.vec("pnorm_log", x), log_ndtr(x)
    assert np.max(np.abs(y - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-13


def test_polya_gamma_moments(oracle):
    for z in (0.0, 0.5, 2.0, 5.0, 12.0):
        w = oracle.pg_draws(11, z, 400000)
        mean = np.tanh(z / 2) / (2 * z) if z > 0 else 0.25
        var = (np.sinh(z) - z) / (4 * z ** 3 * np.cosh(z / 2) ** 2) if z > 0 else 1.0 / 24
        assert abs(w.mean() - mean) < 4.5 * np.sqrt(var / len(w))
        assert abs(w.var() / var - 1) < 0.03
        assert w.min() > 0


def test_c3_unbiased_estimator_matches_long_chain(oracle):
    mdl = cases.c3_model(120, 5, seed=3)
    e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=1), k=30, m=150, lag=1, nrep=1500, nthreads=8)
    assert np.isfinite(e["meetingtime"]).all() and e["meetingtime"].max() < 60
    u = e["uestimator"]
    ch = oracle.sample_coupled_chains(mdl, R.Draws(seed=7), m=20000, lag=1, stride=20100, nrep=1)
    s = ch["samples1"][0, 1000:20001]
    # long-chain standard error via batch means
    bm = s[: (len(s) // 50) * 50].reshape(50, -1, s.shape[1]).mean(1)
    se = np.sqrt(bm.var(0, ddof=1) / 50 + u.var(0) / len(u))
    assert np.all(np.abs(u.mean(0) - s.mean(0)) < 5 * se)


def test_c4_natural_statistic_and_meeting(oracle):
    mdl = cases.c4_model(3, 0.1, 0.2, 0.3)
    e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=2), k=50, m=200, lag=1, nrep=64, nthreads=8)
    assert np.isfinite(e["meetingtime"]).all()
    u = e["uestimator"]
    # natural statistic grows with the inverse temperature; per-site value between 0 and 2
    assert np.all(np.diff(u.mean(0)) > 0)
    assert np.all(u.mean(0) > 0) and np.all(u.mean(0) < 2048)
    # estimator == H_bar of the stored natural statistics
    ch = oracle.sample_coupled_chains(mdl, R.Draws(seed=2), m=200, lag=1, stride=4000, nrep=8)
    np.testing.assert_allclose(oracle.h_bar(ch, A.UBM_H_IDENTITY, k=50, m=200), u[:8], rtol=1e-12)


def test_c5_marginal_likelihood_and_inclusion_probabilities(oracle):
    """inst/check/check_vs_mlikelihood.R:79-95 (agreement of marginal-likelihood formulas) and a long-chain check."""
    import ctypes as C
    mdl = cases.c5_model(100, 30, kappa=0.1, s0=10, snr=2.0)
    e = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=1), k=300, m=1500, lag=1, nrep=600, nthreads=8)
    assert np.isfinite(e["meetingtime"]).all()
    u = e["uestimator"]
    ch = oracle.sample_coupled_chains(mdl, R.Draws(seed=7), m=100000, lag=1, stride=100100, nrep=1)
    s = ch["samples1"][0, 2000:100001]
    se = u.std(0) / np.sqrt(len(u)) + 0.02
    assert np.all(np.abs(u.mean(0) - s.mean(0)) < 5 * se)
    assert np.all(s.mean(0)[:10] > 0.9) and np.all(s.mean(0)[10:] < 0.2)      # the 10 true signals are found
    # states are 0/1 vectors with at most s0 ones
    assert set(np.unique(s)) <= {0.0, 1.0} and s.sum(1).max() <= 10


def test_c3_rejsampler_coupling_is_a_valid_coupling(oracle):
    """w_rejsamplerC (src/logisticregressioncoupling.cpp:42-75): chain 1 follows the same single-chain law as under the
    max coupling (same posterior means from the unbiased estimators), and the chains meet."""
    est = {}
    for mode in ("max", "rej_samp"):
        mdl = cases.c3_model(120, 4, seed=3, mode=mode)
        r = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=11), k=20, m=120, lag=1, nrep=400, nthreads=8,
                                            max_iterations=5000)
        assert np.isfinite(r["meetingtime"]).all()
        est[mode] = (r["uestimator"].mean(0), r["uestimator"].std(0, ddof=1) / np.sqrt(400))
    diff = est["max"][0] - est["rej_samp"][0]
    se = np.sqrt(est["max"][1] ** 2 + est["rej_samp"][1] ** 2)
    assert np.all(np.abs(diff) < 4.5 * se)


def test_c3_crn_beta_step_is_a_valid_coupling(oracle):
    """pgg_coupled_kernel of the vignette (vignettes/polyagammagibbs.Rmd:189-214): chain 1 follows the same single-chain law
    as under sample_beta's maximal coupling (same posterior means from the unbiased estimators), and the chains meet."""
    est = {}
    for bc in ("max", "crn"):
        mdl = cases.c3_model(120, 4, seed=3, beta_coupling=bc)
        r = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=21), k=20, m=120, lag=1, nrep=400, nthreads=8, max_iterations=5000)
        assert np.isfinite(r["meetingtime"]).all()
        est[bc] = (r["uestimator"].mean(0), r["uestimator"].std(0, ddof=1) / np.sqrt(400), r["meetingtime"].mean())
    diff = est["max"][0] - est["crn"][0]
    se = np.sqrt(est["max"][1] ** 2 + est["crn"][1] ** 2)
    assert np.all(np.abs(diff) < 5 * se)
    assert est["crn"][2] < 200


def test_blasso_unbiased_estimators_match_a_long_chain(oracle):
    """get_blasso (R/bayesianlasso.R:22-98): the coupled Gibbs sampler meets, and the unbiased estimators of the posterior
    means of (beta, tau2, sigma2) agree with a long single chain"""
    mdl = cases.blasso_model(120, 8, 1.0)
    r = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=3), k=30, m=200, lag=1, nrep=300, nthreads=8, max_iterations=20000)
    assert np.isfinite(r["meetingtime"]).all() and r["meetingtime"].mean() < 50
    u = r["uestimator"]
    ch = oracle.sample_coupled_chains(mdl, R.Draws(seed=5), m=30000, lag=1, stride=30100, nrep=1, max_iterations=30050)
    s = ch["samples1"][0, 2000:30001]
    bm = s[: (len(s) // 40) * 40].reshape(40, -1, s.shape[1]).mean(1)
    se = np.sqrt(bm.var(0, ddof=1) / 40 + u.var(0, ddof=1) / len(u))
    assert np.all(np.abs(u.mean(0) - s.mean(0)) < 5 * se)
    assert np.all(s[:, 8:] > 0)                                   # tau2 and sigma2 are positive
