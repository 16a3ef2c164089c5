"""Single-launch timing case for the variable-selection kernel (a -DUBM_VS_TIMING build prints per-section cycles).
Usage: python tools/time_case_c5.py [nrep] [k m]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unbiasedmcmc_b200 import workloads as cases
from unbiasedmcmc_b200 import registry as R
from unbiasedmcmc_b200.engine import Engine
eng = Engine(0)
nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 2960
k, m = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (7500, 15000)
mdl = cases.c5_model()
eng.sample_meetingtime(mdl, R.Draws(seed=2), lag=1, nrep=8, max_iterations=100)      # warm-up: Gram matrix, module load
t = time.time()
res = eng.sample_unbiasedestimator(mdl, R.Draws(seed=1), k=k, m=m, lag=1, nrep=nrep, want_bins=False)
dt = time.time() - t
print("nrep", nrep, "wall", dt, "est/s", nrep / dt, "mean tau", res["meetingtime"].mean(), "max", res["meetingtime"].max())
