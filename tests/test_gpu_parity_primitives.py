"""GPU parity: stand-alone coupling / MVN primitives through the C ABI vs the oracle, bit-exact."""
import numpy as np
import pytest

from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R

pytestmark = pytest.mark.gpu


def test_scalar_couplings_bit_exact(engine, oracle):
    for seed, args in ((1, (0.2, -0.8, 0.4, 1.7)), (2, (3.0, 3.0, 1.0, 1.0))):
        g = engine.rnorm_max_coupling(*args, R.Draws(seed=seed), n=5000, rep_offset=11)
        c = oracle.rnorm_max_coupling(*args, R.Draws(seed=seed), n=5000, rep_offset=11)
        assert np.array_equal(g["xy"], c["xy"]) and np.array_equal(g["identical"], c["identical"])
    g = engine.rnorm_reflectionmax(0.2, 1.5, 0.8, R.Draws(seed=3), n=5000)
    c = oracle.rnorm_reflectionmax(0.2, 1.5, 0.8, R.Draws(seed=3), n=5000)
    assert np.array_equal(g["xy"], c["xy"]) and np.array_equal(g["identical"], c["identical"])
    # replay tapes
    rng = np.random.default_rng(0)
    dr = R.Draws(tape_u=rng.uniform(size=(64, 200)), tape_n=rng.standard_normal((64, 200)))
    g = engine.rnorm_max_coupling(0.0, 2.0, 1.0, 0.5, dr, n=64)
    c = oracle.rnorm_max_coupling(0.0, 2.0, 1.0, 0.5, dr, n=64)
    assert np.array_equal(g["xy"], c["xy"])


@pytest.mark.parametrize("d", [2, 7, 100])
def test_mvn_primitives_bit_exact(engine, oracle, d):
    rng = np.random.default_rng(d)
    G = rng.standard_normal((d, d))
    Sigma = G @ G.T / d + np.eye(d)
    U = np.linalg.cholesky(Sigma).T
    Ui = np.triu(np.linalg.inv(U))
    mu1, mu2 = rng.standard_normal(d), rng.standard_normal(d) * 0.1
    n = 300
    g = engine.rmvnorm_reflectionmax(mu1, mu2, U, Ui, R.Draws(seed=9), n=n)
    c = oracle.rmvnorm_reflectionmax(mu1, mu2, U, Ui, R.Draws(seed=9), n=n)
    assert np.array_equal(g["xy"], c["xy"]) and np.array_equal(g["identical"], c["identical"])
    gs = engine.fast_rmvnorm_chol(n, mu1, U, R.Draws(seed=10))
    cs = oracle.fast_rmvnorm_chol(n, mu1, U, R.Draws(seed=10))
    assert np.array_equal(gs, cs)
    assert np.array_equal(engine.fast_dmvnorm_chol_inverse(gs, mu1, Ui), oracle.fast_dmvnorm_chol_inverse(cs, mu1, Ui))
    # identical arguments: norm(z) < 1e-15 branch (src/mvnorm.cpp:200-210)
    g = engine.rmvnorm_reflectionmax(mu1, mu1, U, Ui, R.Draws(seed=9), n=8)
    assert g["identical"].all() and np.array_equal(g["xy"][:, 0], g["xy"][:, 1])


def test_rmvnorm_reflectionmax_rejects_univariate(engine):
    from unbiasedmcmc_b200.engine import UbmError
    with pytest.raises(UbmError) as e:     # R/mvnorm_couplings.R:96-98 stop()
        engine.rmvnorm_reflectionmax([0.0], [1.0], [[1.0]], [[1.0]], R.Draws(seed=1), n=2)
    assert e.value.status == A.UBM_ERR_INVALID


@pytest.mark.parametrize("d", [2, 5, 49])
def test_mvn_max_couplings_bit_exact(engine, oracle, d):
    """rmvnorm_max (src/mvnorm.cpp:91-129) and rmvnorm_max_chol (:133-174): xy, identical and nattempts, bit for bit."""
    rng = np.random.default_rng(100 + d)
    G1, G2 = rng.standard_normal((d, d)), rng.standard_normal((d, d))
    S1 = G1 @ G1.T / d + np.eye(d)
    S2 = 0.8 * (G2 @ G2.T / d + np.eye(d))
    mu1, mu2 = rng.standard_normal(d) * 0.2, rng.standard_normal(d) * 0.2
    U1, U2 = np.linalg.cholesky(S1).T, np.linalg.cholesky(S2).T
    I1, I2 = np.triu(np.linalg.inv(U1)), np.triu(np.linalg.inv(U2))
    n = 400
    g = engine.rmvnorm_max(mu1, mu2, S1, S2, R.Draws(seed=21), n=n, rep_offset=5)
    c = oracle.rmvnorm_max(mu1, mu2, S1, S2, R.Draws(seed=21), n=n, rep_offset=5)
    for key in ("xy", "identical", "nattempts"):
        assert np.array_equal(g[key], c[key]), key
    g = engine.rmvnorm_max_chol(mu1, mu2, U1, U2, I1, I2, R.Draws(seed=22), n=n)
    c = oracle.rmvnorm_max_chol(mu1, mu2, U1, U2, I1, I2, R.Draws(seed=22), n=n)
    for key in ("xy", "identical", "nattempts"):
        assert np.array_equal(g[key], c[key]), key
    assert 0 < g["identical"].mean() < 1 or d > 20
    # replay tapes (rows long enough for the rejection loops; exhaustion is reported identically)
    dr = R.Draws(tape_u=rng.uniform(size=(32, 64)), tape_n=rng.standard_normal((32, 64 * d)))
    g = engine.rmvnorm_max(mu1, mu2, S1, S2, dr, n=32)
    c = oracle.rmvnorm_max(mu1, mu2, S1, S2, dr, n=32)
    assert np.array_equal(g["xy"], c["xy"]) and np.array_equal(g["nattempts"], c["nattempts"])


def test_signed_measure_bit_exact(engine, oracle):
    """ubm_c_chains_to_measure vs the oracle on stored chains of C1 (d = 1) and C2 (d = 7), and the R-named wrapper."""
    from tests import cases
    for mdl, lag, k, m in ((cases.c1_model(), 5, 10, 60), (cases.c2_model(7, "dense", "offset"), 2, 4, 30)):
        ch = oracle.sample_coupled_chains(mdl, R.Draws(seed=3), m=m, lag=lag, stride=3000, nrep=25)
        g = engine.c_chains_to_measure(ch, k=k, m=m)
        c = oracle.c_chains_to_measure(ch, k=k, m=m)
        for a, b in zip(g, c):
            for key in ("atoms", "weights", "MCMC"):
                assert np.array_equal(a[key], b[key]), key
    with pytest.raises(Exception):
        engine.c_chains_to_measure(ch, k=5, m=3)            # k must be <= m


def test_tv_upper_bounds_bit_exact(engine, oracle):
    """L-lag TV upper bounds (vignettes/polyagammagibbs.Rmd:238-244) from C1 meeting times with lag 30."""
    from tests import cases
    from unbiasedmcmc_b200 import api
    mt = engine.sample_meetingtime(cases.c1_model(), R.Draws(seed=5), lag=30, nrep=5000)["meetingtime"]
    t = np.arange(0, 400, 3)
    g, c = engine.tv_upper_bounds(mt, 30, t), oracle.tv_upper_bounds(mt, 30, t)
    assert np.array_equal(g, c)
    assert np.array_equal(g, np.array([api.tv_upper_bound(mt, 30, tj).mean() for tj in t]))
    assert np.all(np.diff(g) <= 0) and g[0] > 1 and g[-1] < 0.2       # slow mixing between the two modes


@pytest.mark.parametrize("d", [1, 3, 40, 128])
def test_covariance_forms_bit_exact(engine, oracle, d):
    """fast_rmvnorm_ / fast_dmvnorm_ (src/mvnorm.cpp:9-26, :49-67; R/mvnorm.R:12,48): LLT on the host, factor forms on the device"""
    rng = np.random.default_rng(d)
    G = rng.standard_normal((d, d + 2))
    S = G @ G.T / d + 0.5 * np.eye(d)
    mean = rng.normal(size=d)
    g = engine.fast_rmvnorm(300, mean, S, R.Draws(seed=5), rep_offset=3)
    c = oracle.fast_rmvnorm(300, mean, S, R.Draws(seed=5), rep_offset=3)
    assert np.array_equal(g, c)
    x = rng.normal(size=(50, d))
    assert np.array_equal(engine.fast_dmvnorm(x, mean, S), oracle.fast_dmvnorm(x, mean, S))
    if d >= 3:
        from scipy import stats
        np.testing.assert_allclose(engine.fast_dmvnorm(x, mean, S), stats.multivariate_normal(mean, S).logpdf(x), rtol=1e-5, atol=1e-4)
        bad = S.copy(); bad[0, 0] = -1.0
        from unbiasedmcmc_b200.engine import UbmError
        with pytest.raises(UbmError):
            engine.fast_rmvnorm(2, mean, bad, R.Draws(seed=1))
