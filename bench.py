#!/usr/bin/env python
"""bench.py — unbiased estimators/sec (completed coupled replicates) on N B200s of one node.

A "step" is one batch of `nrep` independent coupled-chain replicates of the workload run to max(tau, m) with
H_{k:m} accumulated on the device.  Headline workload (BASELINE.json configs[1]): MVN target d = 100, RW-MH with
reflection-maximal coupling (inst/reproducescalingdimension shape), 1e5 replicates per GPU per step (weak scaling).

  value     kernel-resident throughput: model already uploaded, results left in HBM (ubm_plan_*), CUDA events
  e2e       the same batch through the public host-buffer C-ABI call (ubm_sample_unbiasedestimator): model upload
            (H2D), kernels, per-replicate results + summary back to host (D2H), wall clock around the call
  roofline  FP64-FMA roofline of the driver kernel: algorithmic flops (from the realised tau / iteration sums)
            / kernel time, against the FP64 peak measured live on this device (MEASURED_PEAKS.json has no FP64)
  cpu_baseline / --impl reference: the CPU oracle (C++ restatement of the reference; R cannot run here) on all
            host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import cases  # noqa: E402
from unbiasedmcmc_b200 import _abi as A  # noqa: E402
from unbiasedmcmc_b200 import registry as R  # noqa: E402

METRIC = "unbiased estimators/sec (completed coupled replicates)"
UNIT = "estimators/s"


def workload(name):
    """model, kwargs of the estimator driver, replicates per GPU per step, description, flops model"""
    if name == "c2":
        d = 100
        mdl = cases.c2_model(d, "dense", "target")
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=None, k=2000, m=10000, lag=1)
        desc = ("C2: N(0,Sigma) d=100 (Sigma = inverse of one Wishart(d,I) draw), RW-MH proposal Sigma/d, "
                "reflection-max coupling, init from target, lag=1, k=2000 (~q95 of tau), m=5k=10000, h(x)=x")
        gemv = d * (d + 1) / 2 * 2.0   # flops of one triangular vector-matrix product (fma = 2)

        def flops(S, nrep, kw=kw):
            n, sum_tau, sum_iter = S[0] + S[1], S[2], S[5]
            lag = kw["lag"]
            # rinit: 2 proposals + 2 densities; lag single steps: 2 products each; coupled steps: 5; afterwards: 2
            return gemv * (4 * n + 2 * lag * n + 5 * (sum_tau - lag * S[0]) + 2 * (sum_iter - sum_tau))
        return mdl, kw, 100000, desc, flops, "fp64"
    if name == "c1":
        mdl = cases.c1_model()
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=cases.c1_bins(), k=500, m=2000, lag=500)
        desc = ("C1: README bimodal 0.5N(-4,1)+0.5N(4,1), RW-MH sd 2, reflection-max coupling, lag=500, k=500, m=2000, "
                "h(x)=x + 80-bin histogram")

        def flops(S, nrep):
            # ~110 flop-equivalents per target evaluation incl. specials at 20 each; cost = kernel calls
            return 110.0 * S[4]
        return mdl, kw, 1 << 20, desc, flops, "fp64"
    if name == "c4":
        mdl = cases.c4_model()
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=None, k=1000, m=2000, lag=1)
        desc = ("C4: Ising 32x32 periodic, parallel tempering over 16 inverse temperatures 0.30..0.55, pswap 0.01, coupled "
                "systematic-scan Gibbs sweeps (exact wavefront), lag=1, CI-sized horizon k=1000, m=2000 "
                "(run.batch.ising.R uses k=1e5, m=2e5), h = 16 natural statistics")

        def flops(S, nrep):
            # a site update = 1 uniform word (1/4 Philox4x32-10, ~15 integer ops) + ~10 integer ops
            n, sum_tau, sum_iter = S[0] + S[1], S[2], S[5]
            site_updates = 0.99 * 16 * 1024 * (sum_iter + (sum_tau - S[0]))
            return 25.0 * site_updates
        return mdl, kw, 2368, desc, flops, "int32"
    if name == "c3":
        mdl = cases.c3_credit_like_model()
        n, p_ = 1000, 49
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=None, k=110, m=1100, lag=1)
        desc = ("C3: logistic regression n=1000, p=49, synthetic design with the structure of the German-credit model "
                "matrix (intercept, 7 unscaled quantitative columns, 13 factors as dummies with the real level "
                "frequencies; meeting times median 27 / q95 58 against 42 / 75 on the real data), prior N(0, 10 I), "
                "coupled Polya-Gamma Gibbs (per-observation max-coupled PG draws + max-coupled MVN beta), lag=1, "
                "k=110, m=1100, h = beta")

        def flops(S, nrep):
            # single step = X beta (n p fma) + Gram (n p (p+1)/2 fma) + Cholesky/solves (~p^3/2 fma)
            # + n PG draws (~3e2 flop-equivalents each); a coupled step does all of it for both chains (Gram shares
            # the products) plus the Sigma / second-Cholesky work of rmvnorm_max
            n_all, sum_tau, sum_iter = S[0] + S[1], S[2], S[5]
            gram = n * p_ * (p_ + 1) / 2 * 2.0
            single = 2.0 * n * p_ + gram + p_ ** 3 + 300.0 * n
            coupled = 2.0 * (2.0 * n * p_ + gram + 2.0 * p_ ** 3 + 300.0 * n)
            return single * (sum_iter - (sum_tau - S[0])) + coupled * (sum_tau - S[0])
        return mdl, kw, 1184, desc, flops, "fp64"
    if name == "c5":
        mdl = cases.c5_model()
        kw = dict(h_kind=A.UBM_H_IDENTITY, bins=None, k=7500, m=15000, lag=1)
        desc = ("C5: variable selection n=500, p=1000, SNR 1, g=p^3, s0=100, kappa=2, flip/swap MH with coupled index "
                "sampling, lag=1, CI-sized horizon k=7500, m=15000 (clusterscript.kappas.R uses k=75000, m=150000), "
                "h = gamma (1000 inclusion indicators)")

        def flops(S, nrep):
            # per proposal ~ 4 s^2 flop (gather + factor + solve of the s x s block, s ~ 15 here we use the
            # full Cholesky: s^3/3 + 2 s^2), 1 proposal per single step, 2 per coupled step
            n_all, sum_tau, sum_iter = S[0] + S[1], S[2], S[5]
            s = 15.0
            per = s ** 3 / 3.0 + 2.0 * s * s
            return per * (sum_iter + (sum_tau - S[0]))
        return mdl, kw, 8288, desc, flops, "fp64"      # 4 waves of 148 SMs x 14 resident warps
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.stop_flag = False
        self.index = index
        self.th = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


def run_reference(args, mdl, kw, desc):
    """CPU arm: the oracle (port of the reference) with all host threads on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as O
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    O.sample_unbiasedestimator(mdl, R.Draws(seed=999), nrep=cores, nthreads=cores, **kw)
    per_wave = max(time.perf_counter() - t0, 1e-4)
    total_steps = args.steps + args.warmup
    budget = min(120.0, 20.0 * total_steps) / total_steps      # seconds per step
    nrep = int(max(cores, min(100000, cores * max(1.0, budget / per_wave))))
    for w in range(args.warmup):
        O.sample_unbiasedestimator(mdl, R.Draws(seed=1000 + w), nrep=nrep, nthreads=cores, **kw)
    t0 = time.perf_counter()
    for s in range(args.steps):
        O.sample_unbiasedestimator(mdl, R.Draws(seed=2000 + s), nrep=nrep, nthreads=cores, **kw)
    dt = time.perf_counter() - t0
    val = nrep * args.steps / dt
    sample = f"{nrep} replicates per step x {args.steps} steps of the same workload, {cores} std::threads"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "replicates_per_step": nrep},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is an R package (no R/Rcpp/Eigen in this image): timed the C++ oracle restatement, "
                    "which is faster than the interpreted R loop it restates"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--nrep", type=int, default=0, help="replicates per GPU per step (default: the workload's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    mdl, kw, nrep_default, desc, flops_fn, bound = workload(args.workload)
    nrep = args.nrep or nrep_default
    if args.impl == "reference":
        run_reference(args, mdl, kw, desc)
        return

    import torch
    import torch.distributed as dist
    from unbiasedmcmc_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    dimh = mdl.h_dim(kw["h_kind"])
    nbins = kw["bins"].nbins if kw["bins"] is not None else 0
    plan = eng.plan(mdl, max_nrep=nrep, **kw)
    ext_stream = torch.cuda.ExternalStream(plan.stream(), device=torch.device("cuda", local_rank))
    ptr, slen = plan.summary_dev()

    class _Holder:   # zero-copy torch view of the plan's summary buffer for the NCCL all-reduce
        __cuda_array_interface__ = {"shape": (slen,), "typestr": "<f8", "data": (ptr, False), "version": 3}
    summary_t = torch.as_tensor(_Holder(), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(step):
        # replicate indices are global: rank r owns [ (step*world + r) * nrep, ... ), Philox keyed by index
        plan.launch(seed=cases.SEED, nrep=nrep, rep_offset=(step * world + rank) * nrep)

    launches0 = eng.launch_count()
    for w in range(args.warmup):
        one_step(w)
    barrier()
    ext_stream.synchronize()
    peaks = eng.measure_peaks() if rank == 0 else None
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches1 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, flops_total, cost_total = 0.0, 0.0, 0.0
    t0 = time.perf_counter()
    ev0.record(ext_stream)
    for s in range(args.steps):
        one_step(args.warmup + s)
        S = plan.fetch()["summary"]   # waits for the stream; 8*(6+2*dimh+2*nbins) bytes D2H: the step's result
        km, _ = plan.last_ms()
        kernel_ms += km
        flops_total += flops_fn(S, nrep)
        cost_total += S[4]
        if world > 1:
            dist.all_reduce(summary_t)   # the only collective of the path: a few KB of sums over NVLink
    ev1.record(ext_stream)
    ext_stream.synchronize()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    gpu_launches = eng.launch_count() - launches1
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    value = world * nrep * args.steps / (dev_ms_max * 1e-3)

    # ---- e2e: the public host-buffer call, wall clock, same batch size --------------------------------
    e2e_steps = max(1, min(args.steps, 2))
    res = None
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        res = eng.sample_unbiasedestimator(mdl, R.Draws(seed=cases.SEED), nrep=nrep,
                                           rep_offset=((args.warmup + s) * world + rank) * nrep, want_bins=False, **kw)
    barrier()
    e2e_wall = time.perf_counter() - t0
    t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * nrep * e2e_steps / float(t[0])
    # model parameters uploaded by every host-buffer call (what the ubm_model_* struct points at)
    h2d = sum(a.nbytes for a in getattr(mdl, "_flat", [])) if hasattr(mdl, "_flat") else 0
    h2d += sum(v.nbytes for k_, v in vars(mdl).items() if isinstance(v, np.ndarray) and not k_.startswith("chol"))
    if kw["bins"] is not None:
        h2d += kw["bins"].breaks.nbytes
    d2h = nrep * (3 * dimh + 3) * 8 + A.summary_len(dimh, nbins) * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the driver kernel --------------------------------------------------------------------
    achieved = flops_total / (kernel_ms * 1e-3) * 1e-12
    peak = peaks["fp64_tflops"] if bound == "fp64" else peaks["int32_tops"]
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s" if bound == "fp64" else "Top/s",
                "frac": achieved / peak, "traffic": None,
                "kernel": {"c1": "mix_driver_kernel", "c2": "mvn_driver_kernel", "c3": "pg_driver_kernel",
                           "c4": "ising_driver_kernel", "c5": "varsel_driver_kernel"}[args.workload],
                "kernel_share_of_step": kernel_ms / dev_ms,
                "peak_source": "measured live by ubm_measure_peaks (independent DFMA / IMAD chains, 2 ops each); "
                               "MEASURED_PEAKS.json holds no FP64 / INT32 figure; HBM is not the bound (state stays on "
                               "chip, dram traffic of the driver kernel is a few KB per launch, see profiles/)",
                "fp64_tflops_measured": peaks["fp64_tflops"], "int32_tops_measured": peaks["int32_tops"]}
    tj = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tj):
        tr = json.load(open(tj)).get(roofline["kernel"])
        if tr:   # DRAM bytes per launch, scaled from the committed ncu capture (bytes per replicate x replicates)
            roofline["traffic"] = tr["dram_bytes_per_replicate"] * nrep
            roofline["traffic_source"] = tr["source"]
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        roofline["hbm_gbs_measured_peak"] = json.load(open(mp)).get("hbm_gbs")

    # ---- cpu baseline (bounded sample, rank 0 at N = 1 only) ---------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle_py as O
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        O.sample_unbiasedestimator(mdl, R.Draws(seed=999), nrep=cores, nthreads=cores, **kw)
        per_wave = max(time.perf_counter() - t0, 1e-4)
        n_cpu = int(max(cores, min(nrep, cores * max(1.0, 15.0 / per_wave))))
        t0 = time.perf_counter()
        c = O.sample_unbiasedestimator(mdl, R.Draws(seed=cases.SEED), nrep=n_cpu, nthreads=cores,
                                       rep_offset=(args.warmup * world) * nrep, **kw)
        dt = time.perf_counter() - t0
        same = bool(np.array_equal(c["uestimator"], res["uestimator"][:n_cpu])) if e2e_steps == 1 else None
        if e2e_steps > 1:   # res holds the last e2e step: recompute the matching slice for the check
            g = eng.sample_unbiasedestimator(mdl, R.Draws(seed=cases.SEED), nrep=n_cpu,
                                             rep_offset=(args.warmup * world) * nrep, want_bins=False, **kw)
            same = bool(np.array_equal(c["uestimator"], g["uestimator"]))
        cpu = {"value": n_cpu / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {n_cpu} replicates of the first timed step, oracle C++ port, {cores} std::threads",
               "bit_identical_to_gpu": same}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "replicates_per_gpu_per_step": nrep, "k": kw["k"], "m": kw["m"], "lag": kw["lag"],
                       "l2": "no reuse across steps: chain state lives in registers/shared memory; each step draws "
                             "fresh Philox streams, outputs (nrep x dimh doubles) are written once",
                       "timing": "CUDA events on the engine's stream, max over ranks; wall_ms_max also given"},
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "call": "ubm_sample_unbiasedestimator (host buffers, synchronous)"},
            "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "kernel_steps_per_s_rank0": cost_total / (dev_ms * 1e-3)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
