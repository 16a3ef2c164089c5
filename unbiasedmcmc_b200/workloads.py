"""Seeded synthetic inputs for the five BASELINE configs: the workload definitions used by bench.py,
__graft_entry__.smoke() and the tests (tests/cases.py re-exports this module)."""
import numpy as np

from . import _abi as A
from . import registry as R

SEED = 20251019


def c1_model(coupling=A.UBM_COUPLING_REFLECTION_MAX):
    """README.Rmd:69-84: target 0.5 N(-4,1) + 0.5 N(4,1), proposal variance 4, init N(3, 2^2)."""
    if coupling == A.UBM_COUPLING_MAX:   # inst/reproducebimodal/bimodal.run.R:65 get_pb(3, 10, 10)
        return R.NormalMixtureRWMH([0.5, 0.5], [-4.0, 4.0], [1.0, 1.0], 3.0, 10.0, 10.0, coupling)
    return R.NormalMixtureRWMH([0.5, 0.5], [-4.0, 4.0], [1.0, 1.0], 2.0, 3.0, 2.0, coupling)


def c1_bins():
    return R.Bins(0, np.linspace(-8.0, 8.0, 81))


def mvn_sigma(d, kind="dense", seed=21):
    """scaling.rwmh.reflectionmaxcoupling.R:19-27: 'sparse' AR(0.5) or 'dense' inverse of one Wishart(d, I) draw."""
    if kind == "sparse":
        i = np.arange(d)
        return 0.5 ** np.abs(i[:, None] - i[None, :])
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((d, d))
    W = G.T @ G                      # Wishart(d, I)
    S = np.linalg.inv(W)
    return (S + S.T) / 2


def c2_model(d=100, kind="dense", init="target", mean=None):
    """scaling.rwmh.reflectionmaxcoupling.R:12-103 with scale_type = 'optimal' (Sigma_proposal = Sigma_pi / d)."""
    Sigma = mvn_sigma(d, kind)
    mean_pi = np.zeros(d) if mean is None else np.asarray(mean, dtype=np.float64)
    if init == "target":
        return R.MvnRWMH(mean_pi, Sigma, Sigma / d, mean_pi, Sigma)
    return R.MvnRWMH(mean_pi, Sigma, Sigma / d, np.ones(d), np.eye(d))


def c4_model(nchains=16, pswap=0.01, lo=0.3, hi=0.55):
    """inst/reproduceisingmodel/run.batch.ising.R:169-171: betas = seq(0.3, 0.55, length.out = 16), pswap 0.01."""
    return R.IsingPT(np.linspace(lo, hi, nchains), pswap)


def c3_data(n=1000, p=49, seed=SEED):
    """X = [1 | iid N(0,1) columns, standardised], beta* ~ N(0, 0.5^2 I), Y ~ Bernoulli(expit(X beta*))."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p))
    X[:, 1:] = (X[:, 1:] - X[:, 1:].mean(0)) / X[:, 1:].std(0, ddof=1)
    X[:, 0] = 1.0
    beta = 0.5 * rng.standard_normal(p)
    Y = (rng.uniform(size=n) < 1.0 / (1.0 + np.exp(-X @ beta))).astype(np.float64)
    return Y, X


def c3_model(n=1000, p=49, seed=SEED, mode="max", beta_coupling="max"):
    """germancredit.run.R:18-20: prior b = 0, B = 10 I."""
    Y, X = c3_data(n, p, seed)
    return R.PgLogistic(Y, X, np.zeros(p), 10.0 * np.eye(p), mode=mode, beta_coupling=beta_coupling)


# marginal summaries of the German-credit columns (inst/reproducegermancredit/german_credit.csv): (mean, sd, min, max) of
# the 7 quantitative columns and the level frequencies of the 13 factors — the data themselves are not shipped
_CREDIT_QUANT = [(20.9, 12.05, 4, 72), (3271.25, 2821.34, 250, 18424), (2.97, 1.12, 1, 4), (2.84, 1.10, 1, 4),
                 (35.54, 11.35, 19, 75), (1.41, 0.58, 1, 4), (1.16, 0.36, 1, 2)]
_CREDIT_LEVELS = [[.274, .269, .063, .394], [.040, .049, .530, .088, .293],
                  [.234, .103, .181, .280, .012, .022, .050, .009, .097, .012], [.603, .103, .063, .048, .183],
                  [.062, .172, .339, .174, .253], [.050, .310, .548, .092], [.907, .041, .052],
                  [.282, .232, .332, .154], [.139, .047, .814], [.179, .714, .107], [.022, .200, .630, .148],
                  [.596, .404], [.963, .037]]


def c3_credit_like_data(n=1000, seed=SEED):
    """Synthetic design with the STRUCTURE of the German-credit model matrix (germancredit.preparedata.R:2-15): intercept,
    7 unscaled quantitative columns with the real columns' location / scale / support, 13 independent factors with the
    real level frequencies as treatment-coded dummies: p = 49; Y from a logistic model with P(Y = 1) ~ 0.7.  On the real
    data the oracle's meeting times have median 42 and 95 % quantile 75 (lag 1), which is what k = 110 was chosen for."""
    rng = np.random.default_rng(seed)
    cols = [np.ones(n)]
    for mean, sd, lo, hi in _CREDIT_QUANT:
        shape = (mean / sd) ** 2                              # gamma with the column's mean and sd, rounded and clipped
        cols.append(np.clip(np.round(rng.gamma(shape, mean / shape, size=n)), lo, hi))
    for probs in _CREDIT_LEVELS:
        pr = np.asarray(probs) / np.sum(probs)
        lev = rng.choice(len(pr), size=n, p=pr)
        for l in range(1, len(pr)):
            cols.append((lev == l).astype(np.float64))
    X = np.column_stack(cols)
    beta = np.concatenate([[1.2], [-0.03, -1e-4, -0.25, 0.0, 0.012, -0.2, -0.1], rng.normal(0.0, 0.5, X.shape[1] - 8)])
    eta = X @ beta
    eta = eta - np.quantile(eta, 0.3)                         # about 70 % positive responses
    Y = (rng.uniform(size=n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    return Y, X


def c3_credit_like_model(n=1000, seed=SEED):
    """germancredit.run.R:16-19: prior b = 0, B = 10 I on the German-credit-shaped design."""
    Y, X = c3_credit_like_data(n, seed)
    return R.PgLogistic(Y, X, np.zeros(X.shape[1]), 10.0 * np.eye(X.shape[1]))


def c5_data(n=500, p=1000, snr=1.0, seed=SEED):
    """inst/reproducevarselection/varselection.generatedata.run.R:7-18 (independent design, scaled X and Y)."""
    rng = np.random.default_rng(seed)
    beta = np.zeros(p)
    sig = np.array([2, -3, 2, 2, -3, 3, -2, 3, -2, 3], dtype=np.float64)
    beta[:min(10, p)] = snr * np.sqrt(np.log(p) / n) * sig[:min(10, p)]
    X = rng.standard_normal((n, p))
    X = (X - X.mean(0)) / X.std(0, ddof=1)
    Y = X @ beta + rng.standard_normal(n)
    Y = (Y - Y.mean()) / Y.std(ddof=1)
    return Y, X


def c5_model(n=500, p=1000, kappa=2.0, s0=100, snr=1.0, seed=SEED):
    """varselection.differentp.R:22-26,40: g = p^3, s0 = 100, proportion_singleflip = 0.5."""
    Y, X = c5_data(n, p, snr, seed)
    return R.VariableSelection(Y, X, float(p) ** 3, kappa, min(s0, p), 0.5)


# ---------------------------------------------------------------------------------------------------------------------
def blasso_data(n=120, p=8, seed=SEED):
    """a diabetes-like standardised design (inst/reproducebayesianlasso uses lars::diabetes: n = 442, p = 10 / 64): columns
    centred and scaled to unit norm, a sparse signal, Y centred"""
    rng = np.random.default_rng(seed + 17)
    X = rng.standard_normal((n, p))
    X[:, 1::3] += 0.5 * X[:, [0]]
    X = X - X.mean(0)
    X = X / np.sqrt((X ** 2).sum(0))
    beta = np.zeros(p)
    beta[: max(p // 3, 1)] = 40.0 * rng.standard_normal(max(p // 3, 1))
    Y = X @ beta + rng.standard_normal(n)
    return Y - Y.mean(), X


def blasso_model(n=120, p=8, lam=1.0, seed=SEED):
    Y, X = blasso_data(n, p, seed)
    return R.BayesianLasso(Y, X, lam)


# bench workloads: the five BASELINE.json configs at the horizons the reference's scripts use
# ---------------------------------------------------------------------------------------------------------------------
KERNEL_OF = {"c1": "mix_driver_kernel", "c2": "mvn_driver_kernel", "c3": "pg_driver_kernel", "c4": "ising_driver_kernel",
             "c5": "varsel_driver_kernel"}


def bench_spec(name):
    """-> dict(model, kw (estimator driver arguments), nrep (replicates per GPU per step), desc, flops (summary -> algorithmic
    operations of the step), bound ('fp64' | 'int32'))."""
    if name == "c2":
        d = 100
        mdl = c2_model(d, "dense", "target")
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=None, k=2000, m=10000, lag=1)
        desc = ("C2: N(0,Sigma) d=100 (Sigma = inverse of one Wishart(d,I) draw), RW-MH proposal Sigma/d, "
                "reflection-max coupling, init from target, lag=1, k=2000 (~q95 of tau), m=5k=10000, h(x)=x")
        gemv = d * (d + 1) / 2 * 2.0   # flops of one triangular vector-matrix product (fma = 2)

        def flops(S, kw=kw):
            n, sum_tau, sum_iter = S[0] + S[1], S[2], S[5]
            lag = kw["lag"]
            # rinit: 2 proposals + 2 densities; lag single steps: 2 products each; coupled steps: 5; afterwards: 2
            return gemv * (4 * n + 2 * lag * n + 5 * (sum_tau - lag * S[0]) + 2 * (sum_iter - sum_tau))
        return dict(model=mdl, kw=kw, nrep=100000, desc=desc, flops=flops, bound="fp64")
    if name == "c1":
        mdl = c1_model()
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=c1_bins(), k=500, m=2000, lag=500)
        desc = ("C1: README bimodal 0.5N(-4,1)+0.5N(4,1), RW-MH sd 2, reflection-max coupling, lag=500, k=500, m=2000, "
                "h(x)=x + 80-bin histogram")

        def flops(S):
            # ~110 flop-equivalents per target evaluation incl. specials at 20 each; cost = kernel calls
            return 110.0 * S[4]
        return dict(model=mdl, kw=kw, nrep=1 << 20, desc=desc, flops=flops, bound="fp64")
    if name == "c4":
        mdl = c4_model()
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=None, k=100000, m=200000, lag=1)
        desc = ("C4: Ising 32x32 periodic, parallel tempering over 16 inverse temperatures 0.30..0.55, pswap 0.01, coupled "
                "systematic-scan Gibbs sweeps (the reference's in-place scan, reproduced exactly), lag=1, k=1e5, m=2e5 "
                "(run.batch.ising.R:172-173), h = 16 natural statistics")

        def flops(S):
            # a site update = 1 uniform word (1/4 Philox4x32-10, ~15 integer ops) + ~10 integer ops
            sum_tau, sum_iter = S[2], S[5]
            site_updates = 0.99 * 16 * 1024 * (sum_iter + (sum_tau - S[0]))
            return 25.0 * site_updates
        return dict(model=mdl, kw=kw, nrep=4736, desc=desc, flops=flops, bound="int32", cpu_one_wave=True)
    if name == "c3":
        mdl = c3_credit_like_model()
        n, p_ = 1000, 49
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=None, k=110, m=1100, lag=1)
        desc = ("C3: logistic regression n=1000, p=49, synthetic design with the structure of the German-credit model "
                "matrix (intercept, 7 unscaled quantitative columns, 13 factors as dummies with the real level "
                "frequencies; meeting times median 27 / q95 58 against 42 / 75 on the real data), prior N(0, 10 I), "
                "coupled Polya-Gamma Gibbs (per-observation max-coupled PG draws + max-coupled MVN beta), lag=1, "
                "k=110, m=1100 (germancredit.run.R:102-103), h = beta")

        def flops(S):
            # single step = X beta (n p fma) + Gram (n p (p+1)/2 fma) + Cholesky/solves (~p^3/2 fma)
            # + n PG draws (~3e2 flop-equivalents each); a coupled step does all of it for both chains (Gram shares
            # the products) plus the Sigma / second-Cholesky work of rmvnorm_max
            sum_tau, sum_iter = S[2], S[5]
            gram = n * p_ * (p_ + 1) / 2 * 2.0
            single = 2.0 * n * p_ + gram + p_ ** 3 + 300.0 * n
            coupled = 2.0 * (2.0 * n * p_ + gram + 2.0 * p_ ** 3 + 300.0 * n)
            return single * (sum_iter - (sum_tau - S[0])) + coupled * (sum_tau - S[0])
        return dict(model=mdl, kw=kw, nrep=1184, desc=desc, flops=flops, bound="fp64")
    if name == "c5":
        mdl = c5_model()
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=None, k=75000, m=150000, lag=1)
        desc = ("C5: variable selection n=500, p=1000, SNR 1, g=p^3, s0=100, kappa=2, flip/swap MH with coupled index "
                "sampling, lag=1, k=75000, m=150000 (varselection.clusterscript.kappas.R:38-39), h = gamma (1000 inclusion "
                "indicators)")

        def flops(S):
            # per proposal ~ 4 s^2 flop (rank-one update of the s x s factor + solve), s ~ 15;
            # 1 proposal per single step, 2 per coupled step
            sum_tau, sum_iter = S[2], S[5]
            s = 15.0
            return 4.0 * s * s * (sum_iter + (sum_tau - S[0]))
        return dict(model=mdl, kw=kw, nrep=8288, desc=desc, flops=flops, bound="fp64")
    raise ValueError(f"unknown workload {name}")
