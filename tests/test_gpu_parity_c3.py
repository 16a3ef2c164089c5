"""GPU parity, C3 (Polya-Gamma Gibbs for logistic regression): CUDA kernel through the C ABI vs the oracle, bit-exact
under Philox (draws are addressed by (sweep, observation) cell on both sides)."""
import numpy as np
import pytest

from tests import cases
from tests.test_gpu_parity_c1_c2 import assert_same
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R

pytestmark = pytest.mark.gpu
KEYS = ["meetingtime", "iteration", "cost", "uestimator", "mcmcestimator", "correction", "summary"]


@pytest.mark.parametrize("n,p,lag,k,m,h", [(60, 3, 1, 0, 20, A.UBM_H_IDENTITY), (120, 5, 1, 10, 60, A.UBM_H_BOTH),
                                           (333, 17, 2, 5, 30, A.UBM_H_SQUARE), (1000, 49, 1, 3, 12, A.UBM_H_IDENTITY)])
def test_c3_estimator_bit_exact(engine, oracle, n, p, lag, k, m, h):
    mdl = cases.c3_model(n, p, seed=5)
    nrep = 24 if n < 1000 else 8
    kw = dict(h_kind=h, k=k, m=m, lag=lag, nrep=nrep, max_iterations=200 if n < 1000 else 40)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=3), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=3), nthreads=8, **kw)
    assert_same(g, c, KEYS)


def test_c3_meetingtime_and_chains(engine, oracle):
    mdl = cases.c3_model(150, 6, seed=2)
    g = engine.sample_meetingtime(mdl, R.Draws(seed=8), lag=1, nrep=64, rep_offset=5)
    c = oracle.sample_meetingtime(mdl, R.Draws(seed=8), lag=1, nrep=64, rep_offset=5, nthreads=8)
    assert_same(g, c, ["meetingtime"])
    assert np.isfinite(g["meetingtime"]).all()
    kw = dict(m=25, lag=1, stride=400, nrep=12)
    g = engine.sample_coupled_chains(mdl, R.Draws(seed=9), **kw)
    c = oracle.sample_coupled_chains(mdl, R.Draws(seed=9), **kw)
    assert_same(g, c, ["meetingtime", "iteration", "cost"])
    for r in range(12):
        nrow = int(c["iteration"][r]) + 1
        assert np.array_equal(g["samples1"][r, :nrow], c["samples1"][r, :nrow])
        assert np.array_equal(g["samples2"][r, :nrow - 1], c["samples2"][r, :nrow - 1])


def test_c3_rejsampler_coupling_bit_exact(engine, oracle):
    """sample_w(mode = 'rej_samp'): w_rejsamplerC (src/logisticregressioncoupling.cpp:42-75) instead of w_max_couplingC."""
    mdl = cases.c3_model(200, 7, seed=4, mode="rej_samp")
    kw = dict(h_kind=A.UBM_H_BOTH, k=5, m=40, lag=1, nrep=32, max_iterations=400)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=6), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=6), nthreads=8, **kw)
    assert_same(g, c, KEYS)
    assert np.isfinite(g["meetingtime"]).mean() > 0.5
    # the two couplings share the marginal kernel: identical chains until the first coupled step, different meeting times
    g0 = engine.sample_meetingtime(cases.c3_model(200, 7, seed=4), R.Draws(seed=6), lag=1, nrep=32)
    g1 = engine.sample_meetingtime(mdl, R.Draws(seed=6), lag=1, nrep=32)
    assert not np.array_equal(g0["meetingtime"], g1["meetingtime"])


def test_c3_tape_mode_is_refused(engine):
    from unbiasedmcmc_b200.engine import UbmError
    mdl = cases.c3_model(60, 3)
    with pytest.raises(UbmError) as e:
        engine.sample_meetingtime(mdl, R.Draws(tape_u=np.full((2, 8), 0.5), tape_n=np.zeros((2, 8))), nrep=2)
    assert e.value.status == A.UBM_ERR_UNSUPPORTED
