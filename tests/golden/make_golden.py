"""Regenerates tests/golden/oracle_r01.json: small outputs of the oracle under the engine's numerics contract, stored as
hex floats.  They pin OUR canonical arithmetic (Philox streams, summation orders, Cholesky variants) against accidental
drift between rounds (parity with the reference itself is pinned elsewhere: tests/test_r_stream.py against R's printed
README output, tests/test_ref_parity.py against the reference's compiled C++).  Regenerated in round 2: C2 (the
column sums of the triangular products are now one fma chain in k order (what a chain of DMMAs computes), without the
round-1 cut of the long columns), C3 (Gram in FP64-tensor-core order, reciprocal-diagonal triangular solves, m through the inverse factor)
and C4 (Philox sweeps start on a block boundary).  Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import cases                      # noqa: E402
from unbiasedmcmc_b200 import _abi as A      # noqa: E402
from unbiasedmcmc_b200 import registry as R  # noqa: E402


def hexes(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).reshape(-1)]


def golden_cases():
    """name -> (callable(backend) -> dict of arrays); backend is the oracle module or an Engine."""
    Y5, X5 = cases.c5_data(80, 24, snr=2.0)
    c5 = R.VariableSelection(Y5, X5, 24.0 ** 3, 0.5, 8, 0.5)
    rng = np.random.default_rng(4)
    G = rng.standard_normal((4, 4))
    S1 = G @ G.T / 4 + np.eye(4)
    U1 = np.linalg.cholesky(S1).T
    return {
        "c1_estimator": lambda b, kw: b.sample_unbiasedestimator(cases.c1_model(), R.Draws(seed=11), bins=cases.c1_bins(), k=20, m=80, lag=5, nrep=6, **kw),
        "c2_estimator_d9": lambda b, kw: b.sample_unbiasedestimator(cases.c2_model(9, "dense", "offset"), R.Draws(seed=12), h_kind=A.UBM_H_BOTH, k=5, m=40, lag=2, nrep=5, **kw),
        "c3_estimator": lambda b, kw: b.sample_unbiasedestimator(cases.c3_model(90, 4, seed=3), R.Draws(seed=13), k=3, m=25, lag=1, nrep=4, max_iterations=300, **kw),
        "c4_estimator": lambda b, kw: b.sample_unbiasedestimator(cases.c4_model(3, 0.1, 0.2, 0.3), R.Draws(seed=14), k=4, m=30, lag=1, nrep=4, max_iterations=400, **kw),
        "c5_estimator": lambda b, kw: b.sample_unbiasedestimator(c5, R.Draws(seed=15), k=50, m=200, lag=1, nrep=4, max_iterations=2000, **kw),
        "rmvnorm_max": lambda b, kw: b.rmvnorm_max(np.zeros(4), 0.3 * np.ones(4), S1, 0.8 * S1, R.Draws(seed=16), n=6),
        "rmvnorm_reflectionmax": lambda b, kw: b.rmvnorm_reflectionmax(np.zeros(4), 0.3 * np.ones(4), U1, np.triu(np.linalg.inv(U1)), R.Draws(seed=17), n=6),
    }


KEYS = ("meetingtime", "cost", "uestimator", "bins", "xy", "nattempts")


def record(res):
    return {k: hexes(res[k]) for k in KEYS if k in res and res[k] is not None}


if __name__ == "__main__":
    from oracle import oracle_py as O
    out = {name: record(fn(O, {"nthreads": 1} if "estimator" in name else {})) for name, fn in golden_cases().items()}
    with open(os.path.join(ROOT, "tests", "golden", "oracle_r01.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", {k: {kk: len(vv) for kk, vv in v.items()} for k, v in out.items()})
