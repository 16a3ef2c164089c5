import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unbiasedmcmc_b200 import workloads as cases
from unbiasedmcmc_b200 import registry as R
from unbiasedmcmc_b200.engine import Engine
eng = Engine(0)
nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 296
t = time.time()
res = eng.sample_unbiasedestimator(cases.c3_model(), R.Draws(seed=1), k=110, m=1100, lag=1, nrep=nrep, want_bins=False)
dt = time.time() - t
print("nrep", nrep, "wall", dt, "est/s", nrep / dt, "mean tau", res["meetingtime"].mean(), "max", res["meetingtime"].max())
