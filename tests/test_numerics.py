"""CPU: the numerics contract (include/ubm_numerics.h) against independent implementations.
Philox4x32-10 known-answer vectors are the published Random123 kat_vectors; log/exp/log1p against numpy (libm),
qnorm against scipy.special.ndtri."""
import numpy as np
from scipy.special import ndtri


def test_philox_known_answers(oracle):
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _ulp_err(y, ref):
    return np.nanmax(np.abs(y - ref) / np.spacing(np.abs(ref)))


def test_log_exp_log1p_within_one_ulp(oracle):
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(-700, 700, 400000))
    assert _ulp_err(oracle.vec("log", x), np.log(x)) <= 1.0
    x = rng.uniform(0.5, 2.0, 400000)
    assert _ulp_err(oracle.vec("log", x), np.log(x)) <= 1.0
    x = rng.uniform(-745, 709, 400000)
    assert _ulp_err(oracle.vec("exp", x), np.exp(x)) <= 1.0
    x = rng.uniform(-0.99, 5, 400000)
    assert _ulp_err(oracle.vec("log1p", x), np.log1p(x)) <= 1.5


def test_special_values(oracle):
    e = oracle.vec("exp", np.array([0.0, -0.0, -np.inf, np.inf, -800.0, 800.0]))
    assert e[0] == 1.0 and e[1] == 1.0 and e[2] == 0.0 and np.isinf(e[3]) and e[4] == 0.0 and np.isinf(e[5])
    l = oracle.vec("log", np.array([0.0, 1.0, -1.0, np.inf, 5e-324]))
    assert l[0] == -np.inf and l[1] == 0.0 and np.isnan(l[2]) and l[3] == np.inf
    assert abs(l[4] - np.log(5e-324)) < 1e-12


def test_qnorm_matches_ndtri(oracle):
    rng = np.random.default_rng(1)
    p = rng.uniform(0, 1, 400000)
    y, ref = oracle.vec("qnorm", p), ndtri(p)
    assert np.max(np.abs(y - ref)) < 5e-15
    p = 10.0 ** rng.uniform(-300, -1, 100000)    # all three AS241 branches
    y, ref = oracle.vec("qnorm", p), ndtri(p)
    assert np.max(np.abs(y - ref) / np.abs(ref)) < 2e-15
    q = np.arange(1, 1024) / 1024.0                 # 1 - q is exact: the quantile function is odd about 1/2
    assert np.array_equal(oracle.vec("qnorm", q), -oracle.vec("qnorm", 1 - q))


def test_streams_are_open_interval_and_typed(oracle):
    u = oracle.stream(7, 3, 0, 100000)
    assert u.min() > 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.005
    n = oracle.stream(7, 3, 1, 100000)
    assert abs(n.mean()) < 0.02 and abs(n.std() - 1) < 0.02
    e = oracle.stream(7, 3, 2, 100000)
    assert e.min() > 0 and abs(e.mean() - 1) < 0.02
    # different replicate / seed / type give different streams; same triple is reproducible
    assert not np.array_equal(u[:16], oracle.stream(7, 4, 0, 16))
    assert not np.array_equal(u[:16], oracle.stream(8, 3, 0, 16))
    assert np.array_equal(u[:16], oracle.stream(7, 3, 0, 16))
