"""GPU: the README workflow of the reference (README.Rmd:66-121) written against the mirrored R API
(unbiasedmcmc_b200/api.py), with the README's own numbers as statistical anchors (README.md:146,171)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_readme_workflow():
    from unbiasedmcmc_b200 import api
    target = api.NormalMixtureTarget([0.5, 0.5], [-4.0, 4.0], [1.0, 1.0])        # README.Rmd:69-72
    kernels = api.get_mh_kernels(target, 4)                                        # :74
    single_kernel, coupled_kernel = kernels["single_kernel"], kernels["coupled_kernel"]
    rinit = api.normal_rinit(3, 2)                                                 # :80-84
    nrep = 500
    mt = api.sample_meetingtime(single_kernel, coupled_kernel, rinit, nrep=nrep, seed=1)["meetingtime"]    # :89-92
    assert mt.shape == (nrep,) and np.all(mt >= 2) and np.all(np.isfinite(mt))
    chains = api.sample_coupled_chains(single_kernel, coupled_kernel, rinit, m=2000, lag=500, nrep=nrep, seed=1)  # :97-99
    assert len(chains) == nrep
    c0 = chains[0]
    assert c0["samples1"].shape[0] == c0["iteration"] + 1 and c0["samples2"].shape[0] == c0["iteration"] - 500 + 1
    cost = np.array([c["cost"] for c in chains])
    assert abs(cost.mean() - 2072.788) < 5 * cost.std() / np.sqrt(nrep) + 6       # README.md:146
    hist1 = api.histogram_c_chains(chains, 1, k=500, m=2000, nclass=100)           # README.Rmd:102
    assert set(hist1) == {"mids", "breaks", "proportions", "sd", "width"}
    assert abs(hist1["proportions"].sum() - 1.0) < 0.05
    est = api.H_bar(chains, h="identity", k=500, m=2000)[:, 0]                     # README.Rmd:119
    half = 1.96 * est.std(ddof=1) / np.sqrt(nrep)
    assert abs(est.mean()) < 2 * half and 0.10 < half < 0.20                       # README.md:171: 0.0055 +/- 0.1459
    # signed-measure export (R/c_chains_to_measure_as_list.R:19): sum of weights x atoms reproduces H_bar of that chain
    meas = api.c_chains_to_measure_as_list(chains[:3], 500, 2000)
    for M, e in zip(meas, est[:3]):
        assert M["atoms"].shape[1] == 1 and M["MCMC"].sum() == 1501
        np.testing.assert_allclose(np.dot(M["weights"], M["atoms"][:, 0]), e, rtol=1e-10, atol=1e-12)
    # data-frame form with pruning (R/c_chains_to_dataframe.R:22-53, src/prune.cpp:6-45) against a row-by-row restatement
    df = api.c_chains_to_dataframe(chains[:5], 500, 2000)
    np.testing.assert_allclose(np.dot(df["weight"], df["atom"][:, 0]), est[:5].mean(), rtol=1e-9, atol=1e-12)
    assert np.all(np.diff(df["rep"]) >= 0) and set(df["rep"]) == {1, 2, 3, 4, 5}
    raw = api.c_chains_to_dataframe(chains[:5], 500, 2000, prune=False)
    assert len(df["weight"]) < len(raw["weight"])                       # RW-MH rejections repeat atoms
    M = np.column_stack([raw["rep"], raw["MCMC"], raw["weight"], raw["atom"]])
    M = M[np.lexsort([M[:, 3], M[:, 1], M[:, 0]])]
    ref = [M[0].copy()]
    for i in range(1, len(M)):                                         # src/prune.cpp:19-42, literally
        if M[i, 0] == M[i - 1, 0] and np.abs(M[i, 3:] - M[i - 1, 3:]).sum() < 1e-20:
            ref[-1][2] += M[i, 2]
        else:
            ref.append(M[i].copy())
    ref = np.array(ref)
    ref = ref[np.lexsort([ref[:, 1], ref[:, 0]])]
    assert np.array_equal(ref[:, 3], df["atom"][:, 0]) and np.allclose(ref[:, 2], df["weight"], rtol=1e-12, atol=1e-18)
    # k sweep from one upload: every entry equals the single-pair H_bar
    grid = api.H_bar_grid(chains[:20], [100, 500, 900], 2000)
    for i, kk in enumerate((100, 500, 900)):
        assert np.array_equal(grid[i], api.H_bar(chains[:20], h="identity", k=kk, m=2000))
    # on-the-fly estimator: same numbers without storing the chains
    u = api.sample_unbiasedestimator(single_kernel, coupled_kernel, rinit, k=500, m=2000, lag=500, nrep=nrep, seed=1)
    np.testing.assert_allclose(u["uestimator"][:, 0], est, rtol=1e-11, atol=1e-11)
    assert np.array_equal(u["meetingtime"], np.array([c["meetingtime"] for c in chains]))
    # argument guards print and return None, as in R/unbiasedestimator.R:44-51
    assert api.sample_unbiasedestimator(single_kernel, coupled_kernel, rinit, k=5, m=3) is None
    # a plain closure cannot run on the device
    with pytest.raises(TypeError):
        api.get_mh_kernels(lambda x: -0.5 * x * x, 1.0)


def test_registry_models_through_the_api():
    from tests import cases
    from unbiasedmcmc_b200 import api
    ks = api.ising_pt_kernels(np.linspace(0.3, 0.4, 4), 0.05)
    r = api.sample_unbiasedestimator(ks["single_kernel"], ks["coupled_kernel"], ks["rinit"], k=10, m=60, nrep=8, seed=3)
    assert r["uestimator"].shape == (8, 4) and np.isfinite(r["meetingtime"]).all()
    Y, X = cases.c3_data(150, 6, seed=2)
    ks = api.logisticregression_kernels(Y, X, np.zeros(6), 10 * np.eye(6))
    r = api.sample_meetingtime(ks["single_kernel"], ks["coupled_kernel"], ks["rinit"], nrep=16, seed=2)
    assert np.isfinite(r["meetingtime"]).all()
    Y, X = cases.c5_data(100, 30, snr=2.0)
    ks = api.get_variableselection(Y, X, 30.0 ** 3, 0.1, 10, 0.5)
    r = api.sample_unbiasedestimator(ks["single_kernel"], ks["coupled_kernel"], ks["rinit"], k=300, m=1500, nrep=64, seed=1)
    incl = r["uestimator"].mean(0)       # unbiased but noisy with 64 replicates: compare groups, not coordinates
    assert r["uestimator"].shape == (64, 30) and incl[:10].mean() > 0.8 and abs(incl[10:].mean()) < 0.2
    c = api.rnorm_max_coupling(0.0, 1.0, 1.0, 1.0, n=1000, seed=2)
    assert c["xy"].shape == (1000, 2) and c["identical"].dtype == bool


def test_round2_additions_through_the_api():
    """get_blasso, the CRN beta-step, the positive-variable couplings and the covariance-form MVN helpers by their R names"""
    from tests import cases
    from unbiasedmcmc_b200 import api
    Y, X = cases.blasso_data(120, 8)
    ks = api.get_blasso(Y, X, 1.0)
    r = api.sample_unbiasedestimator(ks["gibbs_kernel"], ks["coupled_gibbs_kernel"], ks["rinit"], k=20, m=100, nrep=64, seed=2)
    assert r["uestimator"].shape == (64, 17) and np.isfinite(r["meetingtime"]).all()
    ols = np.linalg.lstsq(X, Y, rcond=None)[0]
    assert np.all(np.abs(r["uestimator"].mean(0)[:2] - ols[:2]) < 0.2 * np.abs(ols[:2]))      # strong signals: little shrinkage
    Yc, Xc = cases.c3_data(150, 6, seed=2)
    ks = api.logisticregression_kernels(Yc, Xc, np.zeros(6), 10 * np.eye(6), beta_coupling="crn")
    assert np.isfinite(api.sample_meetingtime(ks["single_kernel"], ks["coupled_kernel"], ks["rinit"], nrep=16, seed=2)["meetingtime"]).all()
    g = api.rgamma_coupled(2.0, 2.5, 1.0, 1.2, n=2000, seed=1)
    assert g["xy"].shape == (2000, 2) and 0.5 < g["identical"].mean() < 1.0
    ig = api.rinvgaussian_coupled(1.0, 1.1, 2.0, 2.0, n=500, seed=1)
    assert np.all(ig["xy"] > 0)
    S = np.array([[2.0, 0.3], [0.3, 1.0]])
    x = api.fast_rmvnorm(4000, np.array([1.0, -1.0]), S, seed=3)
    assert np.allclose(np.cov(x.T), S, atol=0.15) and np.allclose(x.mean(0), [1.0, -1.0], atol=0.1)
    from scipy import stats
    np.testing.assert_allclose(api.fast_dmvnorm(x[:5], np.array([1.0, -1.0]), S), stats.multivariate_normal([1.0, -1.0], S).logpdf(x[:5]), atol=1e-6)
