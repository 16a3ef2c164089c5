import os, sys, time
sys.path.insert(0, '.')
from unbiasedmcmc_b200 import workloads as cases
from unbiasedmcmc_b200 import registry as R
from unbiasedmcmc_b200.engine import Engine
eng = Engine(0)
mdl = cases.c5_model()
for nrep in (1184, 2368):
    eng.sample_unbiasedestimator(mdl, R.Draws(seed=1), k=2000, m=4000, lag=1, nrep=nrep, want_bins=False)
    t = time.time()
    res = eng.sample_unbiasedestimator(mdl, R.Draws(seed=1), k=20000, m=40000, lag=1, nrep=nrep, want_bins=False)
    dt = time.time() - t
    print(os.environ.get("UBM_LIB_PATH"), "nrep", nrep, "est/s", round(nrep / dt, 1))
