"""GPU parity, C3 (Polya-Gamma Gibbs for logistic regression): CUDA kernel through the C ABI vs the oracle, bit-exact
under Philox (draws addressed by (sweep, observation) cell on both sides) and under replay tapes (the reference's
sequential order on both sides)."""
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


def _tapes(nrep, L, seed):
    rng = np.random.default_rng(seed)
    return R.Draws(tape_u=rng.uniform(size=(nrep, L)), tape_n=rng.standard_normal((nrep, L)),
                   tape_e=rng.exponential(size=(nrep, L)))


@pytest.mark.parametrize("mode", ["max", "rej_samp"])
def test_c3_replay_tapes(engine, oracle, mode):
    """Replay mode in the REFERENCE'S order: one sequential stream per replicate, observation after observation through
    w_max_couplingC / w_rejsamplerC (src/logisticregressioncoupling.cpp:42-113) with their rejection loops, then the beta
    update.  (The oracle's sequential w-coupling is itself checked against the reference's compiled code on the same
    tapes: tests/test_ref_parity.py::test_w_coupling_sequential_order.)"""
    mdl = cases.c3_model(80, 4, seed=3, mode=mode)
    nrep = 6
    draws = _tapes(nrep, 120000, 17)
    g = engine.sample_meetingtime(mdl, draws, lag=1, nrep=nrep, max_iterations=150)
    c = oracle.sample_meetingtime(mdl, draws, lag=1, nrep=nrep, max_iterations=150)
    assert_same(g, c, ["meetingtime"])
    assert np.isfinite(g["meetingtime"]).any()
    kw = dict(h_kind=A.UBM_H_BOTH, k=3, m=25, lag=2, nrep=nrep, max_iterations=150)
    g = engine.sample_unbiasedestimator(mdl, draws, **kw)
    c = oracle.sample_unbiasedestimator(mdl, draws, **kw)
    assert_same(g, c, KEYS)
    kw = dict(m=20, lag=1, stride=160, nrep=nrep)
    g = engine.sample_coupled_chains(mdl, draws, **kw)
    c = oracle.sample_coupled_chains(mdl, draws, **kw)
    assert_same(g, c, ["meetingtime", "iteration", "cost"])
    for r in range(nrep):
        nrow = int(c["iteration"][r]) + 1
        assert np.array_equal(g["samples1"][r, :nrow], c["samples1"][r, :nrow])
        assert np.array_equal(g["samples2"][r, :nrow - 1], c["samples2"][r, :nrow - 1])


def test_c3_replay_differs_from_cell_addressing_only_in_the_draws(engine, oracle):
    """a tape that records, in sequential order, nothing but fresh draws gives a valid run of the same sampler: the two
    addressing modes are two orders of consumption of one algorithm (meeting times have the same law; here: finite)"""
    mdl = cases.c3_model(60, 3, seed=1)
    g = engine.sample_meetingtime(mdl, _tapes(16, 60000, 5), lag=1, nrep=16, max_iterations=200)
    assert np.isfinite(g["meetingtime"]).all() and g["meetingtime"].min() >= 2


def test_c3_tape_exhaustion_is_an_error(engine):
    from unbiasedmcmc_b200.engine import UbmError
    mdl = cases.c3_model(60, 3)
    with pytest.raises(UbmError) as e:
        engine.sample_meetingtime(mdl, _tapes(2, 50, 1), nrep=2)
    assert e.value.status == A.UBM_ERR_TAPE


def test_c3_credit_like_design_bit_exact(engine, oracle):
    """the German-credit-shaped design of the bench workload (unscaled columns up to 18424; workloads.c3_credit_like_model)"""
    mdl = cases.c3_credit_like_model()
    kw = dict(h_kind=A.UBM_H_IDENTITY, k=4, m=16, lag=1, nrep=6, max_iterations=120)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=12), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=12), nthreads=6, **kw)
    assert_same(g, c, KEYS)


@pytest.mark.parametrize("comp", [0, 3])
def test_c3_bins_on_the_fly(engine, oracle, comp):
    """estimator_bin_ (src/estimator_bin.cpp:15-38) of one regression coefficient, accumulated while the chains run"""
    mdl = cases.c3_model(120, 5, seed=2)
    bins = R.Bins(comp, np.linspace(-3.0, 3.0, 41))
    kw = dict(k=5, m=60, lag=1, nrep=24, max_iterations=2000)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=9), bins=bins, **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=9), bins=bins, nthreads=8, **kw)
    assert_same(g, c, KEYS + ["bins"])
    assert abs(g["bins"].sum(1).mean() - 1.0) < 0.2


@pytest.mark.parametrize("lag,h", [(1, A.UBM_H_IDENTITY), (3, A.UBM_H_BOTH)])
def test_c3_common_random_number_beta_step_bit_exact(engine, oracle, lag, h):
    """beta_coupling='crn': the coupled kernel of the vignette (vignettes/polyagammagibbs.Rmd:189-214): max-coupled PG
    variables, ONE Normal increment for both chains' regression coefficients"""
    mdl = cases.c3_model(150, 6, seed=5, beta_coupling="crn")
    kw = dict(h_kind=h, k=5, m=50, lag=lag, nrep=24, max_iterations=3000)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=12), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=12), nthreads=8, **kw)
    assert_same(g, c, KEYS)
    assert np.isfinite(g["meetingtime"]).all()
    ch_g = engine.sample_coupled_chains(mdl, R.Draws(seed=3), m=30, lag=lag, stride=3100, nrep=6, max_iterations=3000)
    ch_c = oracle.sample_coupled_chains(mdl, R.Draws(seed=3), m=30, lag=lag, stride=3100, nrep=6, max_iterations=3000)
    assert_same(ch_g, ch_c, ["meetingtime", "iteration", "cost"])
    for r in range(6):
        n = int(ch_c["iteration"][r]) + 1
        assert np.array_equal(ch_g["samples1"][r, :n], ch_c["samples1"][r, :n])
        assert np.array_equal(ch_g["samples2"][r, :n - lag], ch_c["samples2"][r, :n - lag])


@pytest.mark.parametrize("n,p", [(90, 57), (130, 60), (64, 1), (33, 9)])
def test_c3_edge_sizes_bit_exact(engine, oracle, n, p):
    """p above 56 takes the 8 x 8 tile lists of nine tiles per warp in the Gram build, p = 60 is the largest supported;
    p = 1 and a number of observations that is not a multiple of the 4-observation groups or of the 32-row tiles"""
    mdl = cases.c3_model(n, p, seed=4)
    kw = dict(k=2, m=12, lag=1, nrep=6, max_iterations=400)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=6), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=6), nthreads=8, **kw)
    assert_same(g, c, KEYS)
