"""The remaining scalar maximal couplings: Gamma, inverse Gamma (R/gamma_couplings.R,
R/invgamma_couplings.R) and inverse Gaussian (src/inversegaussian.cpp), through R/get_max_coupling.R.
CPU part: the oracle against the compiled reference (inverse Gaussian: draw for draw from the same tapes) and against
the distributions themselves (marginals, meeting probability = 1 - total variation).  GPU part: bit-exact vs the oracle."""
import os

import numpy as np
import pytest
from scipy import integrate, stats

from unbiasedmcmc_b200 import registry as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tv(pdf1, pdf2, hi):
    return 0.5 * integrate.quad(lambda x: abs(pdf1(x) - pdf2(x)), 0, hi, limit=400, points=np.linspace(0, hi, 50)[1:-1])[0]


def test_oracle_gamma_family_marginals_and_meeting_probability(oracle):
    n = 200000
    for fam, a1, a2, b1, b2 in (("gamma", 2.5, 3.0, 1.5, 2.0), ("gamma", 0.4, 0.7, 1.0, 0.8), ("invgamma", 3.5, 4.0, 2.0, 2.5)):
        fn = oracle.rgamma_coupled if fam == "gamma" else oracle.rinversegamma_coupled
        r = fn(a1, a2, b1, b2, R.Draws(seed=5), n=n)
        x, y = r["xy"][:, 0], r["xy"][:, 1]
        if fam == "gamma":
            d1, d2 = stats.gamma(a1, scale=1 / b1), stats.gamma(a2, scale=1 / b2)
        else:
            d1, d2 = stats.invgamma(a1, scale=b1), stats.invgamma(a2, scale=b2)
        for s, d in ((x, d1), (y, d2)):            # both marginals are exact
            assert stats.kstest(s[:20000], d.cdf).pvalue > 1e-3
            assert abs(s.mean() - d.mean()) < 5 * d.std() / np.sqrt(n) + 1e-3
        assert np.array_equal(r["identical"], x == y)
        tv = _tv(d1.pdf, d2.pdf, max(d1.ppf(0.999999), d2.ppf(0.999999)))
        assert abs(r["identical"].mean() - (1 - tv)) < 5 * np.sqrt(0.25 / n) + 2e-3      # a maximal coupling


def test_oracle_inverse_gaussian_against_the_compiled_reference(oracle):
    """src/inversegaussian.cpp compiled unmodified (oracle/_ref, log redirected to ubm_log): same tapes, same bits"""
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this box and no prebuilt library)")
    L = ref_py.lib("ubmnum")
    rng = np.random.default_rng(3)
    nid = 0
    for mu1, mu2, l1, l2 in ((1.0, 1.3, 2.0, 2.5), (0.5, 0.5, 1.0, 4.0), (3.0, 0.2, 0.7, 0.9)):
        for trial in range(200):
            tn, tu = rng.standard_normal(64), rng.uniform(size=128)
            got = oracle.rinvgaussian_coupled(mu1, mu2, l1, l2, R.Draws(tape_u=tu[None, :], tape_n=tn[None, :]), n=1)
            tape = ref_py.Tape(L, u=tu, n=tn)
            want = np.zeros(2)
            L.ref_rinvgaussian_coupled(mu1, mu2, l1, l2, ref_py._d(want))
            assert not tape.used()[1]
            assert np.array_equal(got["xy"][0], want), (mu1, mu2, trial)
            nid += int(got["identical"][0])
    assert 0 < nid < 600
    tn, tu = rng.standard_normal(50), rng.uniform(size=50)
    got = oracle.rinvgaussian(1, 1.7, 0.9, R.Draws(tape_u=tu[None, :], tape_n=tn[None, :]))
    tape = ref_py.Tape(L, u=tu, n=tn)
    want = np.zeros(1)
    L.ref_rinvgaussian(1, 1.7, 0.9, ref_py._d(want))
    assert got[0] == want[0]


@pytest.mark.gpu
@pytest.mark.parametrize("fam,p", [("gamma", (2.5, 3.0, 1.5, 2.0)), ("gamma", (0.4, 0.7, 1.0, 0.8)), ("invgamma", (3.5, 4.0, 2.0, 2.5)),
                                   ("invgauss", (1.0, 1.3, 2.0, 2.5)), ("invgauss", (3.0, 0.2, 0.7, 0.9))])
def test_gpu_positive_couplings_bit_exact(engine, oracle, fam, p):
    name = {"gamma": "rgamma_coupled", "invgamma": "rinversegamma_coupled", "invgauss": "rinvgaussian_coupled"}[fam]
    g = getattr(engine, name)(*p, R.Draws(seed=9), n=5000, rep_offset=17)
    c = getattr(oracle, name)(*p, R.Draws(seed=9), n=5000, rep_offset=17)
    assert np.array_equal(g["xy"], c["xy"]) and np.array_equal(g["identical"], c["identical"])
    assert 0 < g["identical"].mean() < 1
    plain = {"gamma": "rgamma", "invgamma": "rinversegamma", "invgauss": "rinvgaussian"}[fam]
    assert np.array_equal(getattr(engine, plain)(3000, p[0], p[2], R.Draws(seed=2)), getattr(oracle, plain)(3000, p[0], p[2], R.Draws(seed=2)))
    # replay tapes, and running off them
    rng = np.random.default_rng(1)
    draws = R.Draws(tape_u=rng.uniform(size=(64, 400)), tape_n=rng.standard_normal((64, 400)))
    g = getattr(engine, name)(*p, draws, n=64)
    c = getattr(oracle, name)(*p, draws, n=64)
    assert np.array_equal(g["xy"], c["xy"]) and np.array_equal(g["identical"], c["identical"])
    from unbiasedmcmc_b200.engine import UbmError
    with pytest.raises(UbmError):
        getattr(engine, name)(*p, R.Draws(tape_u=rng.uniform(size=(64, 1)), tape_n=rng.standard_normal((64, 1))), n=64)
