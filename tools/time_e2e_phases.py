"""Where the host-buffer call spends its time: plan creation, launch (+ sync), fetch, and the one-shot call.
Usage: python tools/time_e2e_phases.py c1"""
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
kw = spec["kw"]
eng = Engine(0)
for rep in range(3):
    t0 = time.perf_counter()
    p = eng.plan(spec["model"], kw["h_kind"], kw["bins"], kw["k"], kw["m"], kw["lag"], 0, nrep)
    t1 = time.perf_counter()
    p.launch(1, nrep, 0)
    ms = p.last_ms()
    t2 = time.perf_counter()
    r = p.fetch(per_replicate=True)
    t3 = time.perf_counter()
    p.close()
    t4 = time.perf_counter()
    r = eng.sample_unbiasedestimator(spec["model"], R.Draws(seed=1), nrep=nrep, want_bins=False, **kw)
    t5 = time.perf_counter()
    print(f"plan_create {1e3*(t1-t0):.1f} ms | launch+sync {1e3*(t2-t1):.1f} (device {ms}) | fetch {1e3*(t3-t2):.1f} | destroy {1e3*(t4-t3):.1f} | "
          f"one-shot host-buffer call {1e3*(t5-t4):.1f}", flush=True)
