"""The oracle restatement against the REFERENCE ITSELF: oracle/_ref holds the reference's own, unmodified C++ sources
(/root/reference/src/*.cpp) compiled against stand-in R / Rcpp / Eigen headers (oracle/ref_shim); both sides are fed
identical draws through typed replay tapes.

Bars.  Integer / index / control-flow results (lattices, sums, bin estimators, pair indices, atom tables, identical flags,
numbers of draws consumed): bit-exact.  Continuous results: bit-exact against the 'ubmnum' flavour (the reference's
sources with log / exp taken from include/ubm_numerics.h) wherever the reference fixes the operation order itself, and
within the stated tolerance where the order is Eigen's (products, LLT) or the transcendentals are libm's."""
import numpy as np
import pytest

from tests import cases
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R

ref_py = pytest.importorskip("oracle.ref_py")
pytestmark = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref not built and /root/reference absent")


@pytest.fixture(scope="module")
def ref():
    return ref_py.lib("libm")


@pytest.fixture(scope="module")
def refu():
    return ref_py.lib("ubmnum")


# ---- src/ising.cpp ---------------------------------------------------------------------------------------------
def test_ising_sum_and_sweeps_bit_exact(oracle, ref):
    rng = np.random.default_rng(1)
    for trial in range(6):
        s1 = (2 * rng.integers(0, 2, 1024) - 1).astype(np.int32)
        s2 = (2 * rng.integers(0, 2, 1024) - 1).astype(np.int32)
        assert ref.ref_ising_sum(ref_py._i(s1), 32) == oracle.ising_sum(s1)                     # src/ising.cpp:11-35
        beta = 0.3 + 0.05 * trial
        u = rng.uniform(size=1024)
        o1, _, probas = oracle.ising_sweep(beta, s1, u)                                          # :43-62
        r1 = s1.copy()
        t = ref_py.Tape(ref, u=u)
        ref.ref_ising_gibbs_sweep(ref_py._i(r1), 32, ref_py._d(probas))
        assert t.used() == ((1024, 0, 0), False)
        assert np.array_equal(o1, r1)
        o1, o2, probas = oracle.ising_sweep(beta, s1, u, s2)                                     # :67-92
        r1, r2 = s1.copy(), s2.copy()
        t = ref_py.Tape(ref, u=u)
        ref.ref_ising_coupled_gibbs_sweep(ref_py._i(r1), ref_py._i(r2), 32, ref_py._d(probas))
        assert t.used() == ((1024, 0, 0), False)
        assert np.array_equal(o1, r1) and np.array_equal(o2, r2)
        assert ref.ref_ising_sum(ref_py._i(r2), 32) == oracle.ising_sum(o2)


def test_ising_sweeps_until_the_lattices_meet(oracle, ref):
    """many coupled sweeps in a row at a hot temperature: both implementations coalesce at the same sweep"""
    rng = np.random.default_rng(2)
    s1 = (2 * rng.integers(0, 2, 1024) - 1).astype(np.int32)
    s2 = (2 * rng.integers(0, 2, 1024) - 1).astype(np.int32)
    o1, o2, r1, r2 = s1.copy(), s2.copy(), s1.copy(), s2.copy()
    met = None
    for sweep in range(400):
        u = rng.uniform(size=1024)
        o1, o2, probas = oracle.ising_sweep(0.3, o1, u, o2)
        ref_py.Tape(ref, u=u)
        ref.ref_ising_coupled_gibbs_sweep(ref_py._i(r1), ref_py._i(r2), 32, ref_py._d(probas))
        assert np.array_equal(o1, r1) and np.array_equal(o2, r2)
        if np.array_equal(o1, o2):
            met = sweep
            break
    assert met is not None


# ---- src/estimator_bin.cpp, src/c_chains_to_measure.cpp, src/prune.cpp over chains of the oracle's driver --------------
@pytest.fixture(scope="module")
def chains(oracle):
    mdl = cases.c1_model()
    return oracle.sample_coupled_chains(mdl, R.Draws(seed=31), m=120, lag=15, stride=4096, nrep=40)


def _rows(ch, r):
    n = int(ch["iteration"][r]) + 1
    lag = ch["lag"]
    return (np.ascontiguousarray(ch["samples1"][r, :n]), n, np.ascontiguousarray(ch["samples2"][r, :n - lag]), n - lag)


@pytest.mark.parametrize("k,m", [(0, 120), (30, 120), (60, 60), (100, 110)])
def test_estimator_bin_bit_exact(oracle, ref, chains, k, m):
    bins = R.Bins(0, np.linspace(-8, 8, 33))
    ob = oracle.estimator_bins(chains, bins, k=k, m=m)
    for r in range(40):
        s1, n1, s2, n2 = _rows(chains, r)
        for b in range(bins.nbins):
            v = ref.ref_estimator_bin(ref_py._d(s1), n1, ref_py._d(s2), n2, 1, chains["meetingtime"][r], 1,
                                      bins.breaks[b], bins.breaks[b + 1], k, m, chains["lag"])
            assert v == ob[r, b], (r, b)


def test_estimator_bin_on_the_fly_equals_reference_on_stored_chains(oracle, ref):
    """the streaming accumulation of sample_unbiasedestimator (what the device kernels do) against estimator_bin_ run by
    the reference on the stored trajectories of the same draws"""
    mdl = cases.c1_model()
    bins = R.Bins(0, np.linspace(-8, 8, 41))
    kw = dict(k=40, m=150, lag=20, nrep=30)
    ue = oracle.sample_unbiasedestimator(mdl, R.Draws(seed=8), bins=bins, **kw)
    ch = oracle.sample_coupled_chains(mdl, R.Draws(seed=8), m=150, lag=20, stride=4096, nrep=30)
    assert np.array_equal(ue["meetingtime"], ch["meetingtime"])
    for r in range(30):
        s1, n1, s2, n2 = _rows(ch, r)
        for b in range(bins.nbins):
            v = ref.ref_estimator_bin(ref_py._d(s1), n1, ref_py._d(s2), n2, 1, ch["meetingtime"][r], 1, bins.breaks[b],
                                      bins.breaks[b + 1], 40, 150, 20)
            assert v == ue["bins"][r, b]


@pytest.mark.parametrize("k,m", [(0, 120), (50, 120), (90, 100)])
def test_c_chains_to_measure_bit_exact(oracle, ref, chains, k, m):
    om = oracle.c_chains_to_measure(chains, k=k, m=m)
    for r in range(40):
        s1, n1, s2, n2 = _rows(chains, r)
        cap = len(om[r]["weights"]) + 4
        atoms, w, part = np.zeros((cap, 1)), np.zeros(cap), np.zeros(cap, dtype=np.int32)
        size = ref.ref_c_chains_to_measure(ref_py._d(s1), n1, ref_py._d(s2), n2, 1, chains["meetingtime"][r], k, m, cap,
                                           ref_py._d(atoms), ref_py._d(w), ref_py._i(part))
        assert size == len(om[r]["weights"])
        assert np.array_equal(atoms[:size], om[r]["atoms"])
        assert np.array_equal(w[:size], om[r]["weights"])
        assert np.array_equal(part[:size].astype(bool), om[r]["MCMC"])


def test_prune_measure_bit_exact(ref):
    from unbiasedmcmc_b200 import api
    rng = np.random.default_rng(5)
    for trial in range(5):
        n = 300
        df = np.column_stack([np.sort(rng.integers(1, 4, n)).astype(float), rng.integers(0, 2, n).astype(float),
                              rng.normal(size=n), np.round(rng.normal(size=(n, 2)), 1)])
        keys = [df[:, j] for j in range(df.shape[1] - 1, 2, -1)] + [df[:, 1], df[:, 0]]
        df = np.ascontiguousarray(df[np.lexsort(keys)])
        mine = api.prune_measure_(df)
        out = np.zeros_like(df)
        rows = ref.ref_prune_measure(ref_py._d(df), df.shape[0], df.shape[1], ref_py._d(out))
        assert rows == mine.shape[0] < n
        assert np.array_equal(out[:rows], mine)


# ---- src/sample_pairs01.cpp ------------------------------------------------------------------------------------------
def test_sample_pair01_indices_equal(oracle, ref):
    mdl = cases.c5_model(60, 40, kappa=0.1, s0=10, snr=2.0)
    rng = np.random.default_rng(7)
    for trial in range(3000):
        sel = (rng.uniform(size=40) < rng.uniform(0.1, 0.9)).astype(np.float64)
        if sel.sum() in (0, 40):
            continue
        u = (rng.integers(0, 2**32, 2).astype(np.float64) + 0.5) * 2.0**-32            # the engine's 32-bit uniforms
        out = np.zeros(2, dtype=np.int32)
        ref_py.Tape(ref, u=u)
        ref.ref_sample_pair01(ref_py._d(sel), 40, ref_py._i(out))
        mine = oracle.sample_pair01(mdl, sel, u[0], u[1])
        assert np.array_equal(out - 1, mine)                                             # the reference is 1-based
        assert sel[mine[0]] == 0 and sel[mine[1]] == 1


# ---- src/PolyaGamma.cpp, src/logisticregressioncoupling.cpp (replay order) ---------------------------------------------
def _tapes(seed, n=200000):
    rng = np.random.default_rng(seed)
    return rng.uniform(size=n), rng.standard_normal(n), rng.exponential(size=n)


def test_logcosh_and_xbeta(oracle, ref, refu):
    for x in (-30.0, -2.5, -1e-3, 0.0, 1e-9, 0.7, 3.0, 45.0):
        assert refu.ref_logcosh(x) == oracle.pg_logcosh(x)                               # same transcendentals: same bits
        assert abs(ref.ref_logcosh(x) - oracle.pg_logcosh(x)) <= 4e-16 * max(1.0, abs(x))
    mdl = cases.c3_model(200, 8)
    beta = np.random.default_rng(3).normal(size=8)
    out = np.zeros(200)
    ref.ref_xbeta(ref_py._d(mdl.X), 200, 8, ref_py._d(beta), ref_py._d(out))
    # the reference's a + b * c is two roundings under R's default build; the engine contracts it to one fma
    np.testing.assert_allclose(out, oracle.pg_xbeta(mdl, beta), rtol=0, atol=1e-14)


def test_rpg_devroye_sequential_draws(oracle, ref, refu):
    """PolyaGamma::draw_like_devroye (src/PolyaGamma.cpp:175-226) through rpg_devroye(1, z): same draws consumed in the same
    order, same values"""
    tu, tn, te = _tapes(11)
    z = np.abs(np.random.default_rng(12).normal(scale=3.0, size=4000))
    z[:5] = [0.0, 1e-8, 1.5624, 1.5626, 40.0]                  # both sides of 1 / TRUNC = 1.5625 (rtigauss branches), extremes
    mine, used = oracle.pg_rpg_seq(R.Draws(tape_u=tu[None], tape_n=tn[None], tape_e=te[None]), z)
    for L, exact in ((refu, True), (ref, False)):
        out = np.zeros_like(z)
        t = ref_py.Tape(L, u=tu, n=tn, e=te)
        L.ref_rpg_devroye(ref_py._d(z), z.size, ref_py._d(out))
        cnt, ex = t.used()
        assert not ex and cnt == tuple(int(v) for v in used)
        if exact:
            assert np.array_equal(out, mine)
        else:
            np.testing.assert_allclose(out, mine, rtol=1e-12, atol=0)
    assert used[1] > 0 and used[2] > 0                          # both proposal branches were exercised


@pytest.mark.parametrize("mode", ["max", "rej_samp"])
def test_w_coupling_sequential_order(oracle, ref, refu, mode):
    """w_max_couplingC / w_rejsamplerC (src/logisticregressioncoupling.cpp:42-115), observation after observation on one
    stream — the replay order of the C3 kernel"""
    mdl = cases.c3_model(300, 6, mode=mode)
    rng = np.random.default_rng(21)
    for trial in range(4):
        b1 = rng.normal(size=6)
        b2 = b1 + rng.normal(scale=[0.0, 1e-3, 0.3, 2.0][trial], size=6)
        tu, tn, te = _tapes(30 + trial)
        mine, used = oracle.pg_w_seq(mdl, R.Draws(tape_u=tu[None], tape_n=tn[None], tape_e=te[None]), b1, b2)
        fn = "ref_w_max_coupling" if mode == "max" else "ref_w_rejsampler"
        for L, exact in ((refu, True), (ref, False)):
            w = np.zeros((300, 2))
            t = ref_py.Tape(L, u=tu, n=tn, e=te)
            getattr(L, fn)(ref_py._d(b1), ref_py._d(b2), ref_py._d(mdl.X), 300, 6, ref_py._d(w))
            cnt, ex = t.used()
            # xbeta_ differs by an fma rounding (see test_logcosh_and_xbeta), so z can differ in the last bit: the draws
            # consumed and the coupling pattern must still agree, the values to 1e-12
            assert not ex and cnt == tuple(int(v) for v in used)
            assert np.array_equal(w[:, 0] == w[:, 1], mine[:, 0] == mine[:, 1])
            np.testing.assert_allclose(w, mine, rtol=1e-12, atol=0)
        if trial == 0:
            assert np.all(mine[:, 0] == mine[:, 1])             # equal beta: every pair couples at the first test


# ---- src/mvnorm.cpp, src/logisticregression.cpp, src/vs_mlikelihood.cpp (Eigen stand-in: order of sums differs) -------
def _spd(d, seed):
    rng = np.random.default_rng(seed)
    Q = rng.normal(size=(d, d))
    return Q @ Q.T / d + np.eye(d)


def _close(a, b, tol=1e-10):
    a, b = np.asarray(a), np.asarray(b)
    assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b)))


@pytest.mark.parametrize("d", [2, 7, 100])
def test_mvnorm_sampling_and_density(oracle, ref, d):
    S = _spd(d, d)
    U = np.linalg.cholesky(S).T
    Ui = np.triu(np.linalg.inv(U))
    mean = np.random.default_rng(d).normal(size=d)
    tn = np.random.default_rng(40 + d).standard_normal(d)
    out = np.zeros(d)
    t = ref_py.Tape(ref, n=tn)
    ref.ref_fast_rmvnorm_cholesky(1, d, ref_py._d(mean), ref_py._d(ref_py.cm(U)), ref_py._d(out))     # src/mvnorm.cpp:30-46
    assert t.used() == ((0, d, 0), False)
    mine = oracle.fast_rmvnorm_chol(1, mean, U, R.Draws(tape_n=tn[None]))[0]
    _close(out, mine)
    t = ref_py.Tape(ref, n=tn)
    ref.ref_fast_rmvnorm(1, d, ref_py._d(mean), ref_py._d(ref_py.cm(S)), ref_py._d(out))              # :9-26 (LLT inside)
    _close(out, mine)
    x = np.random.default_rng(50 + d).normal(size=(5, d))
    dens, dens2 = np.zeros(5), np.zeros(5)
    ref.ref_fast_dmvnorm_cholesky_inverse(5, d, ref_py._d(ref_py.cm(x)), ref_py._d(mean), ref_py._d(ref_py.cm(Ui)), ref_py._d(dens))
    ref.ref_fast_dmvnorm(5, d, ref_py._d(ref_py.cm(x)), ref_py._d(mean), ref_py._d(ref_py.cm(S)), ref_py._d(dens2))
    _close(dens, oracle.fast_dmvnorm_chol_inverse(x, mean, Ui))                                       # :71-86
    _close(dens2, dens)                                                                               # :49-67 (check_mvnorm.R:38-46)
    # the covariance forms of the oracle (its own LLT) against the reference's (Eigen's LLT inside)
    _close(out, oracle.fast_rmvnorm(1, mean, S, R.Draws(tape_n=tn[None]))[0])
    _close(dens2, oracle.fast_dmvnorm(x, mean, S))


@pytest.mark.parametrize("d", [2, 7, 100])
def test_rmvnorm_reflection_max_coupling(oracle, ref, d):
    """src/mvnorm.cpp:185-236: same draws consumed, same identical flag, same pair; `identical` => y is a copy of x"""
    S = _spd(d, 3 * d) / d
    U = np.linalg.cholesky(S).T
    Ui = np.triu(np.linalg.inv(U))
    rng = np.random.default_rng(60 + d)
    seen = set()
    for trial in range(40):
        mu1 = rng.normal(size=d)
        mu2 = mu1 + rng.normal(size=d) * [0.0, 0.02, 0.3, 1.0][trial % 4] / np.sqrt(d)
        tn, tu = rng.standard_normal(d), rng.uniform(size=2)
        xy, ident = np.zeros(2 * d), np.zeros(1, dtype=np.int32)
        t = ref_py.Tape(ref, u=tu, n=tn)
        ref.ref_rmvnorm_reflection_max_coupling(d, ref_py._d(mu1), ref_py._d(mu2), ref_py._d(ref_py.cm(U)), ref_py._d(ref_py.cm(Ui)),
                                                ref_py._d(xy), ref_py._i(ident))
        cnt, ex = t.used()
        mine = oracle.rmvnorm_reflectionmax(mu1, mu2, U, Ui, R.Draws(tape_u=tu[None], tape_n=tn[None]))
        assert not ex and cnt == ((0 if trial % 4 == 0 else 1), d, 0)
        assert bool(ident[0]) == bool(mine["identical"][0])
        _close(xy[:d], mine["xy"][0, 0])
        _close(xy[d:], mine["xy"][0, 1])
        if ident[0]:
            assert np.array_equal(xy[:d], xy[d:]) and np.array_equal(mine["xy"][0, 0], mine["xy"][0, 1])
        seen.add((trial % 4, bool(ident[0])))
    assert (0, True) in seen and any(not i for _, i in seen) and any(i and q > 0 for q, i in seen)


@pytest.mark.parametrize("d", [3, 49])
def test_rmvnorm_max_couplings(oracle, ref, d):
    """src/mvnorm.cpp:91-129 (covariances, `<`) and :133-174 (Cholesky factors, `<=`): draws consumed, flags, attempts, pairs"""
    rng = np.random.default_rng(70 + d)
    S1, S2 = _spd(d, 1) / d, _spd(d, 2) / d
    U1, U2 = np.linalg.cholesky(S1).T, np.linalg.cholesky(S2).T
    I1, I2 = np.triu(np.linalg.inv(U1)), np.triu(np.linalg.inv(U2))
    flags = set()
    for trial in range(30):
        mu1 = rng.normal(size=d)
        mu2 = mu1 + rng.normal(size=d) * [0.0, 0.05, 0.4][trial % 3] / np.sqrt(d)
        tn, tu = rng.standard_normal(60 * d), rng.uniform(size=60)
        xy, ident, na = np.zeros(2 * d), np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        t = ref_py.Tape(ref, u=tu, n=tn)
        ref.ref_rmvnorm_max_coupling(d, ref_py._d(mu1), ref_py._d(mu2), ref_py._d(ref_py.cm(S1)), ref_py._d(ref_py.cm(S2)),
                                     ref_py._d(xy), ref_py._i(ident), ref_py._i(na))
        cnt, ex = t.used()
        mine = oracle.rmvnorm_max(mu1, mu2, S1, S2, R.Draws(tape_u=tu[None], tape_n=tn[None]))
        assert not ex and cnt == (1 + int(na[0]), d * (1 + int(na[0])), 0)
        assert bool(ident[0]) == bool(mine["identical"][0]) and int(na[0]) == int(mine["nattempts"][0])
        _close(xy[:d], mine["xy"][0, 0]); _close(xy[d:], mine["xy"][0, 1])
        flags.add(bool(ident[0]))
        t = ref_py.Tape(ref, u=tu, n=tn)
        ref.ref_rmvnorm_max_coupling_cholesky(d, ref_py._d(mu1), ref_py._d(mu2), ref_py._d(ref_py.cm(U1)), ref_py._d(ref_py.cm(U2)),
                                              ref_py._d(ref_py.cm(I1)), ref_py._d(ref_py.cm(I2)), ref_py._d(xy), ref_py._i(ident))
        cnt, ex = t.used()
        mine = oracle.rmvnorm_max_chol(mu1, mu2, U1, U2, I1, I2, R.Draws(tape_u=tu[None], tape_n=tn[None]))
        assert not ex and cnt == (1 + int(mine["nattempts"][0]), d * (1 + int(mine["nattempts"][0])), 0)
        assert bool(ident[0]) == bool(mine["identical"][0])
        _close(xy[:d], mine["xy"][0, 0]); _close(xy[d:], mine["xy"][0, 1])
    assert flags == {True, False}


def test_m_sigma_function(oracle, ref):
    """src/logisticregression.cpp:32-56 (triple loop + LLT + solve + inverse) and sigma_ (:5-23)"""
    mdl = cases.c3_model(400, 12)
    om = np.random.default_rng(80).gamma(2.0, 0.1, size=400)
    mine = oracle.pg_m_sigma(mdl, om)
    p = 12
    m, A_, L, Li = np.zeros(p), np.zeros(p * p), np.zeros(p * p), np.zeros(p * p)
    ref.ref_m_sigma_function(400, p, ref_py._d(om), ref_py._d(mdl.X), ref_py._d(ref_py.cm(mine["invB"])), ref_py._d(mine["ktk"]),
                             ref_py._d(m), ref_py._d(A_), ref_py._d(L), ref_py._d(Li))
    cmj = lambda a: a.reshape(p, p).T
    _close(m, mine["m"])
    _close(np.tril(cmj(L)), np.tril(mine["Cholesky_inverse"]))
    _close(np.tril(cmj(Li)), np.tril(mine["Cholesky"]))
    X = mdl.X.reshape(p, 400).T
    S = np.zeros(p * p)
    ref.ref_sigma(400, p, ref_py._d(mdl.X), ref_py._d(om), ref_py._d(S))
    _close(cmj(S) + mine["invB"], cmj(A_))
    _close(cmj(A_), X.T @ (om[:, None] * X) + mine["invB"])


def test_marginal_likelihood_float32_reference_bound(oracle, ref):
    """src/vs_mlikelihood.cpp:5-28 computes in float32 (Eigen .inverse()); the engine in FP64 from G = X'X (DESIGN: documented
    deviation).  This bounds the difference on the log marginal likelihood over selections of every size the sampler visits;
    differences of log-ml between neighbouring selections (what the MH ratio uses) are bounded the same way."""
    mdl = cases.c5_model(500, 200, kappa=2.0, s0=100, snr=1.0)
    rng = np.random.default_rng(90)
    X = mdl.X.reshape(200, 500).T
    Y2 = float(mdl.Y @ mdl.Y)
    worst = 0.0
    for s in (0, 1, 2, 5, 10, 20, 40, 80):
        for trial in range(4):
            sel = np.zeros(200)
            sel[rng.choice(200, s, replace=False)] = 1.0
            r = ref.ref_marginal_likelihood_c_2(500, 200, ref_py._d(sel), ref_py._d(mdl.X), ref_py._d(mdl.Y), Y2, mdl.g)
            o = oracle.varsel_ml(mdl, sel)
            Xs = X[:, sel > 0]
            l = float(mdl.Y @ Xs @ np.linalg.solve(Xs.T @ Xs, Xs.T @ mdl.Y)) if s else 0.0
            exact = -(s + 1) / 2 * np.log(mdl.g + 1) - 500 / 2 * np.log(Y2 - mdl.g / (mdl.g + 1) * l)
            assert abs(o - exact) <= 1e-9 * abs(exact)            # the FP64 path is the accurate one
            worst = max(worst, abs(r - o) / abs(o))
    assert worst < 2e-4, worst                                    # float32 evaluation: ~1e-5 relative on log-ml


def test_blasso_conditional_against_the_compiled_reference(oracle, refu):
    """blassoconditional (src/blassoutil.cpp:17-41, Eigen .inverse() + fast_rmvnorm_'s LLT) against the oracle's Cholesky route,
    same normal tape: the regression draw, the residual sum of squares and beta' D^-1 beta agree to rounding"""
    from unbiasedmcmc_b200 import workloads as W
    for n, p in ((60, 5), (150, 12)):
        mdl = W.blasso_model(n, p, 1.3)
        Y, X = W.blasso_data(n, p)
        rng = np.random.default_rng(p)
        tau2, sigma2, tn = rng.uniform(0.5, 2.0, p), 1.7, rng.standard_normal(p)
        beta, nb = np.zeros(p), np.zeros(2)
        t = ref_py.Tape(refu, n=tn)
        refu.ref_blassoconditional(n, p, ref_py._d(ref_py.f64(Y)), ref_py._d(ref_py.cm(X)), ref_py._d(ref_py.f64(X.T @ Y)),
                                   ref_py._d(ref_py.cm(X.T @ X)), ref_py._d(ref_py.f64(tau2)), sigma2, ref_py._d(beta), ref_py._d(nb))
        assert t.used() == ((0, p, 0), False)
        b, norm, bdb = oracle.blasso_conditional(mdl, R.Draws(tape_n=tn[None]), tau2, sigma2)
        _close(beta, b, 1e-9)
        _close(np.array([norm, bdb]), nb, 1e-9)
