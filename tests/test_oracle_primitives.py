"""CPU: the oracle's stand-alone coupling / MVN primitives against the analytic facts the reference's check scripts use
(inst/check/check_normalcouplings.R:33-35, check_reflmaxcoupling.R:28-30, check_mvnorm.R:38-46)."""
import numpy as np
from scipy import stats

from unbiasedmcmc_b200 import registry as R


def _tv_overlap_normal(mu1, mu2, s1, s2):
    x = np.linspace(min(mu1, mu2) - 12 * max(s1, s2), max(mu1, mu2) + 12 * max(s1, s2), 400001)
    return np.trapezoid(np.minimum(stats.norm.pdf(x, mu1, s1), stats.norm.pdf(x, mu2, s2)), x)


def test_rnorm_max_coupling_marginals_and_meeting_probability(oracle):
    mu1, mu2, s1, s2, n = 0.2, -0.8, 0.4, 1.7, 200000
    r = oracle.rnorm_max_coupling(mu1, mu2, s1, s2, R.Draws(seed=3), n=n)
    x, y, ident = r["xy"][:, 0], r["xy"][:, 1], r["identical"]
    assert stats.kstest((x - mu1) / s1, "norm").pvalue > 1e-3
    assert stats.kstest((y - mu2) / s2, "norm").pvalue > 1e-3
    p = _tv_overlap_normal(mu1, mu2, s1, s2)
    assert abs(ident.mean() - p) < 4 * np.sqrt(p * (1 - p) / n)
    assert np.array_equal(ident, x == y)              # identical flag <=> bitwise equality


def test_rnorm_reflectionmax_marginals_and_meeting_probability(oracle):
    mu1, mu2, s, n = 0.2, 1.5, 0.8, 200000
    r = oracle.rnorm_reflectionmax(mu1, mu2, s, R.Draws(seed=4), n=n)
    x, y, ident = r["xy"][:, 0], r["xy"][:, 1], r["identical"]
    assert stats.kstest((x - mu1) / s, "norm").pvalue > 1e-3
    assert stats.kstest((y - mu2) / s, "norm").pvalue > 1e-3
    p = _tv_overlap_normal(mu1, mu2, s, s)
    assert abs(ident.mean() - p) < 4 * np.sqrt(p * (1 - p) / n)
    # the raw function returns y = mu2 + sigma (xdot + z) when identical: equal to x up to rounding only
    # (get_mh_kernels copies proposal1 in that case, R/get_mh_kernels.R:52-53)
    assert np.max(np.abs(x[ident] - y[ident])) < 1e-12


def test_rmvnorm_reflectionmax_and_densities(oracle):
    d, n = 4, 100000
    rng = np.random.default_rng(0)
    G = rng.standard_normal((d, d))
    Sigma = G @ G.T + d * np.eye(d)
    U = np.linalg.cholesky(Sigma).T                  # R's chol()
    Ui = np.triu(np.linalg.inv(U))
    mu1, mu2 = rng.standard_normal(d), rng.standard_normal(d) * 0.5
    r = oracle.rmvnorm_reflectionmax(mu1, mu2, U, Ui, R.Draws(seed=5), n=n)
    x, y, ident = r["xy"][:, 0, :], r["xy"][:, 1, :], r["identical"]
    assert np.array_equal(ident, np.all(x == y, axis=1))            # check_reflmaxcoupling.R:28-30
    for v, mu in ((x, mu1), (y, mu2)):
        assert np.all(np.abs(v.mean(0) - mu) < 5 * np.sqrt(np.diag(Sigma) / n))
        assert np.max(np.abs(np.cov(v.T) - Sigma)) < 0.02 * np.max(np.diag(Sigma)) + 0.1
    # meeting probability of the reflection-maximal coupling = overlap of N(0,1) and N(|z|,1), z = Ui^T (mu1 - mu2)
    z = np.linalg.norm((mu1 - mu2) @ Ui)
    p = 2 * stats.norm.cdf(-z / 2)
    assert abs(ident.mean() - p) < 4 * np.sqrt(p * (1 - p) / n)
    # fast_rmvnorm_chol and fast_dmvnorm_chol_inverse (check_mvnorm.R:38-46: equals the textbook log-density up to the
    # truncated constant 0.9189385 per dimension, src/mvnorm.cpp:75)
    s = oracle.fast_rmvnorm_chol(50000, mu1, U, R.Draws(seed=6))
    assert np.all(np.abs(s.mean(0) - mu1) < 5 * np.sqrt(np.diag(Sigma) / 50000))
    lp = oracle.fast_dmvnorm_chol_inverse(s[:200], mu1, Ui)
    ref = stats.multivariate_normal(mu1, Sigma).logpdf(s[:200])
    np.testing.assert_allclose(lp, ref + d * (0.918938533204672741780329736406 - 0.9189385), rtol=0, atol=1e-10)


def _mvn_pair(d, seed):
    rng = np.random.default_rng(seed)
    G1, G2 = rng.standard_normal((d, d)), rng.standard_normal((d, d))
    S1 = G1 @ G1.T / d + np.eye(d)
    S2 = 0.7 * (G2 @ G2.T / d + np.eye(d))
    mu1, mu2 = rng.standard_normal(d) * 0.3, rng.standard_normal(d) * 0.3
    return mu1, mu2, S1, S2


def _mvn_overlap_mc(mu1, mu2, S1, S2, n=400000, seed=1):
    """integral of min(p, q) by importance sampling from p."""
    rng = np.random.default_rng(seed)
    x = rng.multivariate_normal(mu1, S1, size=n)
    lp = stats.multivariate_normal(mu1, S1).logpdf(x)
    lq = stats.multivariate_normal(mu2, S2).logpdf(x)
    w = np.minimum(1.0, np.exp(lq - lp))
    return w.mean(), w.std() / np.sqrt(n)


def test_rmvnorm_max_marginals_and_meeting_probability(oracle):
    """src/mvnorm.cpp:91-129 and :133-174: marginals N(mu1, S1), N(mu2, S2); P(identical) = overlap of the two densities;
    identical <=> bitwise equality; the two parametrisations agree in law."""
    d, n = 3, 100000
    mu1, mu2, S1, S2 = _mvn_pair(d, 7)
    p, pse = _mvn_overlap_mc(mu1, mu2, S1, S2)
    U1, U2 = np.linalg.cholesky(S1).T, np.linalg.cholesky(S2).T
    for r in (oracle.rmvnorm_max(mu1, mu2, S1, S2, R.Draws(seed=8), n=n),
              oracle.rmvnorm_max_chol(mu1, mu2, U1, U2, np.triu(np.linalg.inv(U1)), np.triu(np.linalg.inv(U2)),
                                      R.Draws(seed=9), n=n)):
        x, y, ident = r["xy"][:, 0, :], r["xy"][:, 1, :], r["identical"]
        assert np.array_equal(ident, np.all(x == y, axis=1))
        for v, mu, S in ((x, mu1, S1), (y, mu2, S2)):
            assert np.all(np.abs(v.mean(0) - mu) < 5 * np.sqrt(np.diag(S) / n))
            assert np.max(np.abs(np.cov(v.T) - S)) < 0.03 * np.max(np.diag(S)) + 0.02
        assert abs(ident.mean() - p) < 4 * np.sqrt(p * (1 - p) / n) + 4 * pse
        assert np.all(r["nattempts"][ident] == 0) and np.all(r["nattempts"][~ident] >= 1)
    # equal Normals always meet at the first test
    r = oracle.rmvnorm_max(mu1, mu1, S1, S1, R.Draws(seed=3), n=64)
    assert r["identical"].all()
