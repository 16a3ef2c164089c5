"""GPU parity, C1 (Normal mixture) and C2 (MVN): the CUDA engine, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bar: meeting times, iterations, costs and bin numerators bit-exact;
FP64 estimators bit-exact too (stated tolerance of north_star: 1e-12 relative; we assert equality and report
the max deviation if it ever fails)."""
import numpy as np
import pytest

from tests import cases
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R

pytestmark = pytest.mark.gpu


def assert_same(res_gpu, res_cpu, keys, tol=0.0):
    for k in keys:
        a, b = res_gpu[k], res_cpu[k]
        if a is None and b is None:
            continue
        if tol == 0.0:
            if not np.array_equal(a, b):
                bad = np.argwhere(~(np.asarray(a) == np.asarray(b)))
                dev = np.nanmax(np.abs(np.asarray(a, dtype=float) - np.asarray(b, dtype=float)))
                raise AssertionError(f"{k}: {len(bad)} mismatches, first at {bad[:3].tolist()}, max abs dev {dev}")
        else:
            np.testing.assert_allclose(a, b, rtol=tol, atol=0)


EST_KEYS = ["meetingtime", "iteration", "cost", "uestimator", "mcmcestimator", "correction", "bins", "summary"]


@pytest.mark.parametrize("coupling", [A.UBM_COUPLING_REFLECTION_MAX, A.UBM_COUPLING_MAX])
def test_c1_estimator_philox_bit_exact(engine, oracle, coupling):
    mdl = cases.c1_model(coupling)
    kw = dict(h_kind=A.UBM_H_BOTH, bins=cases.c1_bins(), k=500, m=2000, lag=500, nrep=1500, rep_offset=7)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=11), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=11), nthreads=8, **kw)
    assert_same(g, c, EST_KEYS)
    assert np.all(np.isfinite(g["meetingtime"]))


@pytest.mark.parametrize("lag,k,m", [(1, 0, 1), (1, 0, 50), (10, 10, 50), (3, 7, 7), (50, 0, 50)])
def test_c1_estimator_small_horizons(engine, oracle, lag, k, m):
    mdl = cases.c1_model()
    kw = dict(h_kind=A.UBM_H_IDENTITY, bins=cases.c1_bins(), k=k, m=m, lag=lag, nrep=257)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=5), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=5), **kw)
    assert_same(g, c, EST_KEYS)


def test_c1_meetingtime_and_censoring(engine, oracle):
    mdl = cases.c1_model()
    g = engine.sample_meetingtime(mdl, R.Draws(seed=3), lag=1, nrep=4000)
    c = oracle.sample_meetingtime(mdl, R.Draws(seed=3), lag=1, nrep=4000, nthreads=8)
    assert_same(g, c, ["meetingtime"])
    # max_iterations: censored replicates report +Inf (R/meetingtime.R:35-36)
    g = engine.sample_meetingtime(mdl, R.Draws(seed=3), lag=1, max_iterations=20, nrep=4000)
    c = oracle.sample_meetingtime(mdl, R.Draws(seed=3), lag=1, max_iterations=20, nrep=4000, nthreads=8)
    assert_same(g, c, ["meetingtime"])
    assert np.isinf(g["meetingtime"]).any() and np.isfinite(g["meetingtime"]).any()


def test_c1_estimator_censored_summary(engine, oracle):
    mdl = cases.c1_model()
    kw = dict(bins=cases.c1_bins(), k=5, m=20, lag=2, max_iterations=25, nrep=999)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=9), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=9), **kw)
    assert_same(g, c, EST_KEYS)
    assert g["summary"][1] > 0   # some censored


def test_c1_replay_tapes_bit_exact(engine, oracle):
    """Replay mode: both sides consume the same recorded draws (typed tapes), not Philox."""
    mdl = cases.c1_model()
    nrep, L = 64, 4000
    rng = np.random.default_rng(123)
    draws = R.Draws(tape_u=rng.uniform(size=(nrep, L)), tape_n=rng.standard_normal((nrep, L)))
    kw = dict(bins=cases.c1_bins(), k=100, m=400, lag=50, nrep=nrep)
    g = engine.sample_unbiasedestimator(mdl, draws, **kw)
    c = oracle.sample_unbiasedestimator(mdl, draws, **kw)
    assert_same(g, c, EST_KEYS)
    # a tape recorded from the Philox streams replays the Philox run exactly
    tu = np.stack([oracle.stream(77, r, 0, L) for r in range(nrep)])
    tn = np.stack([oracle.stream(77, r, 1, L) for r in range(nrep)])
    g_t = engine.sample_unbiasedestimator(mdl, R.Draws(tape_u=tu, tape_n=tn), **kw)
    g_p = engine.sample_unbiasedestimator(mdl, R.Draws(seed=77), **kw)
    assert_same(g_t, g_p, EST_KEYS)


def test_c1_tape_exhaustion_is_an_error(engine):
    from unbiasedmcmc_b200.engine import UbmError
    mdl = cases.c1_model()
    rng = np.random.default_rng(1)
    draws = R.Draws(tape_u=rng.uniform(size=(8, 10)), tape_n=rng.standard_normal((8, 10)))
    with pytest.raises(UbmError) as e:
        engine.sample_unbiasedestimator(mdl, draws, k=10, m=100, lag=1, nrep=8)
    assert e.value.status == A.UBM_ERR_TAPE


def test_c1_chains_hbar_bins(engine, oracle):
    mdl = cases.c1_model()
    kw = dict(m=300, lag=20, stride=2048, nrep=200)
    g = engine.sample_coupled_chains(mdl, R.Draws(seed=4), **kw)
    c = oracle.sample_coupled_chains(mdl, R.Draws(seed=4), **kw)
    assert_same(g, c, ["meetingtime", "iteration", "cost"])
    for r in range(200):
        n = int(c["iteration"][r]) + 1
        assert np.array_equal(g["samples1"][r, :n], c["samples1"][r, :n])
        assert np.array_equal(g["samples2"][r, :n - 20], c["samples2"][r, :n - 20])
    hb_g = engine.h_bar(g, A.UBM_H_BOTH, k=50, m=300)
    hb_c = oracle.h_bar(c, A.UBM_H_BOTH, k=50, m=300)
    assert np.array_equal(hb_g, hb_c)
    bins = cases.c1_bins()
    assert np.array_equal(engine.estimator_bins(g, bins, k=50, m=300), oracle.estimator_bins(c, bins, k=50, m=300))
    # on-the-fly estimator == H_bar of the stored chains (same draws), up to summation order
    e = engine.sample_unbiasedestimator(mdl, R.Draws(seed=4), h_kind=A.UBM_H_BOTH, bins=bins, k=50, m=300, lag=20, nrep=200)
    np.testing.assert_allclose(e["uestimator"], hb_g, rtol=1e-12, atol=1e-12)
    assert np.array_equal(e["bins"], engine.estimator_bins(g, bins, k=50, m=300))


def test_c1_split_invariance(engine):
    """Results depend only on (seed, global replicate index), not on how replicates are split over calls."""
    mdl = cases.c1_model()
    kw = dict(k=20, m=100, lag=5)
    a = engine.sample_unbiasedestimator(mdl, R.Draws(seed=8), nrep=300, **kw)
    b1 = engine.sample_unbiasedestimator(mdl, R.Draws(seed=8), nrep=100, rep_offset=0, **kw)
    b2 = engine.sample_unbiasedestimator(mdl, R.Draws(seed=8), nrep=200, rep_offset=100, **kw)
    assert np.array_equal(a["uestimator"], np.concatenate([b1["uestimator"], b2["uestimator"]]))
    assert np.array_equal(a["meetingtime"], np.concatenate([b1["meetingtime"], b2["meetingtime"]]))


@pytest.mark.parametrize("d,kind,init", [(2, "dense", "target"), (5, "sparse", "offset"), (23, "dense", "offset"),
                                         (100, "dense", "target")])
def test_c2_meetingtime_bit_exact(engine, oracle, d, kind, init):
    mdl = cases.c2_model(d, kind, init)
    nrep = 300 if d < 100 else 120
    g = engine.sample_meetingtime(mdl, R.Draws(seed=21), lag=1, nrep=nrep, rep_offset=3)
    c = oracle.sample_meetingtime(mdl, R.Draws(seed=21), lag=1, nrep=nrep, rep_offset=3, nthreads=8)
    assert_same(g, c, ["meetingtime"])


@pytest.mark.parametrize("d,lag,k,m,h", [(5, 1, 0, 30, A.UBM_H_BOTH), (12, 4, 10, 60, A.UBM_H_IDENTITY),
                                         (100, 1, 50, 250, A.UBM_H_SQUARE)])
def test_c2_estimator_bit_exact(engine, oracle, d, lag, k, m, h):
    mean = np.linspace(-1, 1, d) if d == 12 else None     # non-zero target mean exercises the centred path
    mdl = cases.c2_model(d, "dense", "offset", mean=mean)
    nrep = 100 if d < 100 else 40
    kw = dict(h_kind=h, k=k, m=m, lag=lag, nrep=nrep)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=2), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=2), nthreads=8, **kw)
    assert_same(g, c, ["meetingtime", "iteration", "cost", "uestimator", "mcmcestimator", "correction", "summary"])


def test_c2_full_size_properties(engine):
    """BASELINE.json configs[1] at full size (d = 100, k = 2000, m = 10000, lag = 1), where the oracle is too slow to be the
    checker: size-independent properties instead — cost / iteration identities (R/coupled_chains.R:82), uestimator =
    mcmc + correction (R/unbiasedestimator.R:98-102), unbiasedness for E[x] = 0 and E[x^2] = diag(Sigma), the summary's
    fixed-order means, and independence of the results from how the replicates are split over launches."""
    d, k, m, nrep = 100, 2000, 10000, 1184
    mdl = cases.c2_model(d)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=4), h_kind=A.UBM_H_BOTH, k=k, m=m, lag=1, nrep=nrep)
    tau, it, cost = g["meetingtime"], g["iteration"], g["cost"]
    assert np.isfinite(tau).all() and tau.min() >= 2
    assert np.array_equal(it, np.maximum(tau, m).astype(np.int64))
    assert np.array_equal(cost, 1 + 2 * (tau - 1) + np.maximum(0, it - tau))
    u = g["uestimator"]          # (mcmc + correction) / (m - k + 1): one rounding away from the sum of the two outputs
    np.testing.assert_allclose(u, g["mcmcestimator"] + g["correction"], rtol=1e-14, atol=1e-15)
    se = u.std(0, ddof=1) / np.sqrt(nrep)
    truth = np.concatenate([np.zeros(d), np.diag(cases.mvn_sigma(d))])
    zs = (u.mean(0) - truth) / se
    assert np.abs(zs).max() < 5.0 and (np.abs(zs) > 3).sum() <= 4
    assert g["summary"][0] == nrep and g["summary"][1] == 0            # all done, none censored
    np.testing.assert_allclose(g["summary"][6:6 + 2 * d] / nrep, u.mean(0), rtol=1e-10, atol=1e-13)
    h1 = engine.sample_unbiasedestimator(mdl, R.Draws(seed=4), h_kind=A.UBM_H_BOTH, k=k, m=m, lag=1, nrep=300)
    h2 = engine.sample_unbiasedestimator(mdl, R.Draws(seed=4), h_kind=A.UBM_H_BOTH, k=k, m=m, lag=1, nrep=nrep - 300,
                                         rep_offset=300)
    assert np.array_equal(u, np.concatenate([h1["uestimator"], h2["uestimator"]]))
    assert np.array_equal(tau, np.concatenate([h1["meetingtime"], h2["meetingtime"]]))


def test_c2_censoring_and_ragged_batches(engine, oracle):
    """max_iterations reached before meeting (R/meetingtime.R:41-44: meetingtime = Inf), batches that do not fill the 16
    slots of a CTA, and more CTAs than replicates: identical to the oracle, censored replicates counted in the summary."""
    mdl = cases.c2_model(12, "dense", "offset")
    for nrep in (1, 7, 17, 150):
        kw = dict(k=5, m=40, lag=2, nrep=nrep, max_iterations=200)      # about 60 % of the replicates are censored
        g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=12), **kw)
        c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=12), nthreads=4, **kw)
        assert_same(g, c, ["meetingtime", "iteration", "cost", "uestimator", "mcmcestimator", "correction", "summary"])
        assert g["summary"][1] == np.isinf(g["meetingtime"]).sum()
    assert np.isinf(g["meetingtime"]).any() and np.isfinite(g["meetingtime"]).any()
    assert np.all(g["iteration"][np.isinf(g["meetingtime"])] == 200)
    g = engine.sample_meetingtime(mdl, R.Draws(seed=12), lag=3, nrep=33, max_iterations=25)
    c = oracle.sample_meetingtime(mdl, R.Draws(seed=12), lag=3, nrep=33, max_iterations=25, nthreads=4)
    assert_same(g, c, ["meetingtime"])


def test_c2_chains_and_hbar(engine, oracle):
    mdl = cases.c2_model(7, "dense", "offset")
    kw = dict(m=40, lag=3, stride=1024, nrep=50)
    g = engine.sample_coupled_chains(mdl, R.Draws(seed=6), **kw)
    c = oracle.sample_coupled_chains(mdl, R.Draws(seed=6), **kw)
    assert_same(g, c, ["meetingtime", "iteration", "cost"])
    for r in range(50):
        n = int(c["iteration"][r]) + 1
        assert np.array_equal(g["samples1"][r, :n], c["samples1"][r, :n])
        assert np.array_equal(g["samples2"][r, :n - 3], c["samples2"][r, :n - 3])
    assert np.array_equal(engine.h_bar(g, A.UBM_H_IDENTITY, k=5, m=40), oracle.h_bar(c, A.UBM_H_IDENTITY, k=5, m=40))


def test_c2_replay_tapes(engine, oracle):
    d = 6
    mdl = cases.c2_model(d, "sparse", "offset")
    nrep, steps = 16, 400
    rng = np.random.default_rng(5)
    draws = R.Draws(tape_u=rng.uniform(size=(nrep, 2 * steps)), tape_n=rng.standard_normal((nrep, d * (steps + 2))))
    g = engine.sample_meetingtime(mdl, draws, lag=2, max_iterations=steps - 5, nrep=nrep)
    c = oracle.sample_meetingtime(mdl, draws, lag=2, max_iterations=steps - 5, nrep=nrep)
    assert_same(g, c, ["meetingtime"])


def test_invalid_arguments(engine):
    from unbiasedmcmc_b200.engine import UbmError
    mdl = cases.c1_model()
    for kw in (dict(k=5, m=3, lag=1), dict(k=0, m=3, lag=4), dict(k=0, m=3, lag=0)):   # R/unbiasedestimator.R:44-51
        with pytest.raises(UbmError) as e:
            engine.sample_unbiasedestimator(mdl, R.Draws(seed=1), nrep=4, **kw)
        assert e.value.status == A.UBM_ERR_INVALID


def test_c2_exact_length_tapes(engine, oracle):
    """A tape that holds exactly the draws a replicate consumes is enough: the producer warps' prefetch past its end is
    not an error (one draw fewer is)."""
    from unbiasedmcmc_b200.engine import UbmError
    d = 12
    mdl = cases.c2_model(d)
    rng = np.random.default_rng(5)
    big = R.Draws(tape_u=rng.uniform(size=(1, 4000)), tape_n=rng.standard_normal((1, 40000)))
    kw = dict(k=2, m=30, lag=1, nrep=1)
    c = oracle.sample_unbiasedestimator(mdl, big, **kw)
    it, tau = int(c["iteration"][0]), int(c["meetingtime"][0])
    # normals: 2 d (rinit) + d per kernel call; uniforms: 1 per single step, 2 per coupled step (R/get_mh_kernels.R:32-57)
    n_used = 2 * d + d * it
    u_used = 1 + 2 * (tau - 1) + (it - tau)
    exact = R.Draws(tape_u=big.tapes[0][:, :u_used].copy(), tape_n=big.tapes[1][:, :n_used].copy())
    c2 = oracle.sample_unbiasedestimator(mdl, exact, **kw)
    g = engine.sample_unbiasedestimator(mdl, exact, **kw)
    assert_same(c2, c, EST_KEYS)
    assert_same(g, c, EST_KEYS)
    for short in (R.Draws(tape_u=exact.tapes[0][:, :-1].copy(), tape_n=exact.tapes[1]),
                  R.Draws(tape_u=exact.tapes[0], tape_n=exact.tapes[1][:, :-1].copy())):
        with pytest.raises(UbmError) as e:
            engine.sample_unbiasedestimator(mdl, short, **kw)
        assert e.value.status == A.UBM_ERR_TAPE


def test_stored_chain_estimators_reject_censored_replicates(engine):
    """H_bar / estimator_bin on a chain stopped by max_iterations (meetingtime = Inf): R's H_bar stops with an error
    (R/H_bar.R:45); here UBM_ERR_INVALID instead of reading past the stored rows"""
    from unbiasedmcmc_b200.engine import UbmError
    mdl = cases.c1_model()
    ch = engine.sample_coupled_chains(mdl, R.Draws(seed=3), m=10, lag=1, max_iterations=12, stride=64, nrep=400)
    assert np.isinf(ch["meetingtime"]).any()
    for call in (lambda: engine.h_bar(ch, A.UBM_H_IDENTITY, k=2, m=10),
                 lambda: engine.h_bar_grid(ch, [1, 2], [10, 10], A.UBM_H_IDENTITY),
                 lambda: engine.estimator_bins(ch, cases.c1_bins(), k=2, m=10)):
        with pytest.raises(UbmError) as e:
            call()
        assert e.value.status == A.UBM_ERR_INVALID
    ok = {k: (v[np.isfinite(ch["meetingtime"])] if isinstance(v, np.ndarray) else v) for k, v in ch.items()}
    ok = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in ok.items()}
    assert np.all(np.isfinite(engine.h_bar(ok, A.UBM_H_IDENTITY, k=2, m=10)))
    with pytest.raises(UbmError):
        engine.h_bar(ok, A.UBM_H_IDENTITY, k=11, m=10)          # k > m


@pytest.mark.parametrize("d,comp", [(7, 0), (20, 13), (100, 99)])
def test_c2_on_the_fly_bins_bit_exact(engine, oracle, d, comp):
    """estimator_bin_ (src/estimator_bin.cpp:10-42) accumulated while the chains run — C2: integer numerators through
    global reductions, finalised per replicate; the same values as the reference's routine on stored chains."""
    mdl = cases.c2_model(d)
    bins = R.Bins(comp, np.linspace(-3.0, 3.0, 25) * np.sqrt(cases.mvn_sigma(d)[comp, comp]))
    nrep = 40 if d < 100 else 10
    kw = dict(h_kind=A.UBM_H_IDENTITY, bins=bins, k=10, m=60, lag=3, nrep=nrep, rep_offset=3)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=21), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=21), nthreads=8, **kw)
    assert_same(g, c, EST_KEYS)
    assert g["bins"].sum() > 0
    stride = int(g["iteration"].max()) + 2
    ch = engine.sample_coupled_chains(mdl, R.Draws(seed=21), m=60, lag=3, stride=stride, nrep=nrep, rep_offset=3)
    stored = engine.estimator_bins(ch, bins, k=10, m=60)
    assert np.array_equal(stored, g["bins"])


@pytest.mark.parametrize("d", [3, 9, 31, 128])
def test_c2_edge_dimensions_bit_exact(engine, oracle, d):
    """dimensions that are not multiples of the 4-wide k steps / 8-wide output tiles of the tensor-core phases, and the
    largest supported one"""
    mdl = cases.c2_model(d, "dense", "offset")
    kw = dict(h_kind=A.UBM_H_IDENTITY, k=3, m=25, lag=2, nrep=40 if d < 128 else 24, max_iterations=4000)
    g = engine.sample_unbiasedestimator(mdl, R.Draws(seed=8), **kw)
    c = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=8), nthreads=8, **kw)
    assert_same(g, c, ["meetingtime", "iteration", "cost", "uestimator", "mcmcestimator", "correction", "summary"])
