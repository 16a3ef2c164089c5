"""Small single-launch cases for ncu (one driver-kernel launch each). Usage: python tools/prof_case.py c1|c2|c3|c4|c5 [nrep]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unbiasedmcmc_b200 import workloads as cases
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R
from unbiasedmcmc_b200.engine import Engine

which = sys.argv[1]
eng = Engine(0)
if which == "c2":
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 16
    res = eng.sample_unbiasedestimator(cases.c2_model(100), R.Draws(seed=1), k=400, m=2000, lag=1, nrep=nrep, want_bins=False)
elif which == "c2bench":      # the bench workload's k, m on one wave of replicates
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 16
    res = eng.sample_unbiasedestimator(cases.c2_model(100), R.Draws(seed=1), k=2000, m=10000, lag=1, nrep=nrep, want_bins=False)
elif which == "c4":
    # keep it tiny: under `ncu --set full` (about 40 replays) 296 replicates with m = 400 did not finish in 10 minutes
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
    res = eng.sample_unbiasedestimator(cases.c4_model(), R.Draws(seed=1), k=20, m=40, lag=1, nrep=nrep, want_bins=False,
                                       max_iterations=80)
elif which == "c2meet":
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 16
    res = eng.sample_meetingtime(cases.c2_model(100), R.Draws(seed=1), lag=1, nrep=nrep)
elif which == "c1":
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
    res = eng.sample_unbiasedestimator(cases.c1_model(), R.Draws(seed=1), bins=cases.c1_bins(), k=500, m=2000, lag=500,
                                       nrep=nrep, want_bins=False, per_replicate=False)
elif which == "c3":
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 148
    res = eng.sample_unbiasedestimator(cases.c3_model(), R.Draws(seed=1), k=10, m=60, lag=1, nrep=nrep, want_bins=False)
if which != "c5":
    print(which, "ok", res["meetingtime"][:4] if res.get("meetingtime") is not None else res["summary"][:6])
if which == "c5":
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 16
    res = eng.sample_unbiasedestimator(cases.c5_model(), R.Draws(seed=1), k=1500, m=3000, lag=1, nrep=nrep, want_bins=False)
    print("c5 ok", res["meetingtime"][:4])
