"""The oracle pinned to REAL R OUTPUT.

oracle/r_stream.hpp emulates the R runtime's default stream (set.seed -> Mersenne-Twister -> unif_rand; inversion
norm_rand; Ahrens-Dieter exp_rand).  Run on that stream, the oracle's restatement of the reference's R drivers and kernels
(R/meetingtime.R:24-48, R/coupled_chains.R:39-87, R/get_mh_kernels.R:20-79, R/rnorm_reflectionmax.R:18-39, R/H_bar.R:19-54)
must reproduce the literal outputs of the reference's README (README.md:146 and :171), which were printed by R itself
after set.seed(1).  These are the only literal outputs in the reference tree."""
import numpy as np
import pytest

from tests import cases


def test_r_stream_published_first_draws(oracle):
    # the first draws after set.seed() that every R user has seen
    s = oracle.RSession(1)
    np.testing.assert_allclose(s.runif(3), [0.2655087, 0.3721239, 0.5728534], atol=5e-8)
    s = oracle.RSession(1)
    np.testing.assert_allclose(s.rnorm(3), [-0.6264538, 0.1836433, -0.8356286], atol=5e-8)
    s = oracle.RSession(1)
    np.testing.assert_allclose(s.rexp(3), [0.7551818, 1.1816428, 0.1457067], atol=5e-8)
    s = oracle.RSession(42)
    np.testing.assert_allclose(s.runif(2), [0.9148060, 0.9370754], atol=5e-8)
    s = oracle.RSession(123)
    np.testing.assert_allclose(s.rnorm(3), [-0.56047565, -0.23017749, 1.55870831], atol=5e-9)


def readme_session(oracle, record=False):
    """README.Rmd:66-121 on the emulated stream: set.seed(1); 500 x sample_meetingtime; 500 x sample_coupled_chains(m =
    2000, lag = 500); H_bar(k = 500, m = 2000)."""
    mdl = cases.c1_model()
    s = oracle.RSession(1)
    mt = oracle.sample_meetingtime(mdl, s, lag=1, nrep=500)["meetingtime"]
    if record:
        s.record = (4400, 2300, 0)
    ch = oracle.sample_coupled_chains(mdl, s, m=2000, lag=500, stride=2101, nrep=500)
    return mdl, mt, ch, s


def test_readme_literals_reproduced(oracle):
    mdl, mt, ch, s = readme_session(oracle)
    assert np.all(np.isfinite(mt)) and np.all(np.isfinite(ch["meetingtime"]))
    assert ch["iteration"].max() <= 2100
    # README.md:146   mean(sapply(coupledchains, function(x) x$cost))  ->  [1] 2072.788
    assert f"{ch['cost'].mean():.7g}" == "2072.788"
    # README.md:171   95% confidence interval for the mean: 0.005471002 +/- 0.1458675
    est = oracle.h_bar(ch, k=500, m=2000)[:, 0]
    assert f"{est.mean():.7g}" == "0.005471002"
    assert f"{1.96 * est.std(ddof=1) / np.sqrt(len(est)):.7g}" == "0.1458675"


def test_recorded_tapes_replay_the_session(oracle):
    """The typed tapes recorded from the R session replay to the same chains through UBM_DRAWS_TAPE (the form in which
    the device consumes 'the reference's own recorded draws')."""
    from unbiasedmcmc_b200 import registry as R
    mdl, mt, ch, s = readme_session(oracle, record=True)
    tu, tn, _ = s.tapes
    used_u, used_n = np.isfinite(tu).sum(1), np.isfinite(tn).sum(1)
    assert used_u.max() < tu.shape[1] and used_n.max() < tn.shape[1]
    rp = oracle.sample_coupled_chains(mdl, R.Draws(tape_u=np.nan_to_num(tu), tape_n=np.nan_to_num(tn)), m=2000, lag=500,
                                      stride=2101, nrep=500)
    for key in ("meetingtime", "iteration", "cost"):
        assert np.array_equal(rp[key], ch[key])
    assert np.array_equal(rp["samples1"], ch["samples1"], equal_nan=True)
    assert np.array_equal(rp["samples2"], ch["samples2"], equal_nan=True)
