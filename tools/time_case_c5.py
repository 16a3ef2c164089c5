"""Steady-state timing case for the variable-selection kernel (the debug timing build prints per-phase cycles of one warp).
Usage: UBM_LIB_PATH=unbiasedmcmc_b200/_lib/libubm_timing.so python tools/time_case_c5.py [nrep] [k m]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unbiasedmcmc_b200 import workloads as cases
from unbiasedmcmc_b200 import registry as R
from unbiasedmcmc_b200.engine import Engine

eng = Engine(0)
nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 16
k, m = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (20000, 40000)
t = time.time()
res = eng.sample_unbiasedestimator(cases.c5_model(), R.Draws(seed=1), k=k, m=m, lag=1, nrep=nrep, want_bins=False)
dt = time.time() - t
print("nrep", nrep, "wall", dt, "est/s", nrep / dt, "mean tau", res["meetingtime"].mean())
