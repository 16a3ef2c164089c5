"""Single-launch wall-clock timing of a bench workload at a chosen replicate count / horizon.
Usage: python tools/time_workload.py c4 [nrep] [k m]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unbiasedmcmc_b200 import workloads as W
from unbiasedmcmc_b200 import registry as R
from unbiasedmcmc_b200.engine import Engine

name = sys.argv[1]
spec = W.bench_spec(name)
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else spec["nrep"]
kw = dict(spec["kw"])
if len(sys.argv) > 4:
    kw["k"], kw["m"] = int(sys.argv[3]), int(sys.argv[4])
eng = Engine(0)
eng.sample_unbiasedestimator(spec["model"], R.Draws(seed=2), nrep=min(nrep, 64), want_bins=False,
                             **{**kw, "k": kw["lag"] + 1, "m": kw["lag"] + 8})   # warm-up
t = time.time()
res = eng.sample_unbiasedestimator(spec["model"], R.Draws(seed=1), nrep=nrep, want_bins=False, **kw)
dt = time.time() - t
print(name, "nrep", nrep, "k", kw["k"], "m", kw["m"], "wall %.3f" % dt, "est/s %.1f" % (nrep / dt), "mean tau %.1f" % res["meetingtime"].mean(), flush=True)
