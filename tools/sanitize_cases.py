"""Tiny runs of every driver kernel (for compute-sanitizer where it is available — it is closed on the GPU pool this was
developed on — or as a quick all-kernels smoke).  Usage:
compute-sanitizer --tool memcheck python tools/sanitize_cases.py [c1 c2 c3 c4 c5 blasso]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from unbiasedmcmc_b200 import workloads as W, registry as R, _abi as A
from unbiasedmcmc_b200.engine import Engine

eng = Engine(0)
which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5", "blasso"]
cases = {
    "c1": (W.c1_model(), dict(bins=W.c1_bins(), k=5, m=30, lag=5, nrep=64)),
    "c2": (W.c2_model(12, "dense", "offset"), dict(k=3, m=40, lag=1, nrep=40, bins=R.Bins(3, np.linspace(-4, 4, 9)))),
    "c3": (W.c3_model(70, 9, seed=4), dict(k=2, m=8, lag=1, nrep=3, max_iterations=60)),
    "c4": (W.c4_model(5, 0.3, 0.3, 0.4), dict(k=2, m=6, lag=1, nrep=9, max_iterations=12)),
    "c5": (W.c5_model(60, 20, kappa=0.1, s0=6, snr=2.0), dict(k=5, m=40, lag=1, nrep=6, max_iterations=400)),
    "blasso": (W.blasso_model(40, 4, 1.0), dict(k=2, m=10, lag=1, nrep=3, max_iterations=200)),
}
for name in which:
    mdl, kw = cases[name]
    r = eng.sample_unbiasedestimator(mdl, R.Draws(seed=5), **kw)
    print(name, "ok", r["meetingtime"][:4], flush=True)
