#!/usr/bin/env python
"""bench.py — unbiased estimators/sec (completed coupled replicates) on N B200s of one node.

A "step" is one batch of `nrep` independent coupled-chain replicates of a workload run to max(tau, m) with H_{k:m}
accumulated on the device.  Headline workload (BASELINE.json configs[1]): MVN target d = 100, RW-MH with
reflection-maximal coupling (inst/reproducescalingdimension shape), 1e5 replicates per GPU per step (weak scaling).
The same JSON line carries a `per_workload` block with all five BASELINE configs at the horizons of the reference's
scripts (C1 README bimodal, C3 PG logistic, C4 Ising k=1e5 m=2e5, C5 variable selection k=75000 m=150000), each with
value / e2e / roofline / cpu_baseline, and, for N > 1, the strong-scaling figure of C2 (1e5 replicates in total).

  value     kernel-resident throughput: model already uploaded, results left in HBM (ubm_plan_*), CUDA events on the
            engine's stream; for N > 1 the NCCL all-reduce of the summary buffer is enqueued on that stream, inside the
            timed region, and its result is checked against the per-rank summaries (`nccl_check`)
  e2e       the same batch through the public host-buffer C-ABI call (ubm_sample_unbiasedestimator): model upload
            (H2D), kernels, per-replicate results + summary back to host (D2H), wall clock around the call
  roofline  FP64-FMA (INT32 for C4) roofline of the driver kernel: algorithmic operations (from the realised tau /
            iteration sums) / kernel time, against the peak measured live on this device (MEASURED_PEAKS.json has no
            FP64 / INT32 figure)
  cpu_baseline / --impl reference: the CPU oracle (C++ restatement of the reference, pinned to R's own README output and
            to the reference's compiled C++; the R drivers themselves cannot run here) on all host cores, on a bounded
            sample of the same workload.
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

from unbiasedmcmc_b200 import _abi as A  # noqa: E402
from unbiasedmcmc_b200 import registry as R  # noqa: E402
from unbiasedmcmc_b200 import workloads as W  # noqa: E402
from unbiasedmcmc_b200.multigpu import split_replicates  # noqa: E402

METRIC = "unbiased estimators/sec (completed coupled replicates)"
UNIT = "estimators/s"


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


def cpu_sample(spec, seconds, rep_offset=0, seed=W.SEED, nrep_cap=None):
    """The oracle on all host cores on a bounded sample: -> (estimators/s, result, nrep, cores)."""
    from oracle import oracle_py as O
    cores = os.cpu_count() or 1
    mdl, kw = spec["model"], spec["kw"]
    if spec.get("cpu_one_wave"):     # a single replicate takes most of a minute per core: one replicate per core, no probe
        n_cpu = cores
    else:
        t0 = time.perf_counter()
        O.sample_unbiasedestimator(mdl, R.Draws(seed=999), nrep=cores, nthreads=cores, **kw)
        per_wave = max(time.perf_counter() - t0, 1e-4)
        n_cpu = int(max(cores, min(nrep_cap or spec["nrep"], cores * max(1.0, seconds / per_wave))))
    t0 = time.perf_counter()
    c = O.sample_unbiasedestimator(mdl, R.Draws(seed=seed), nrep=n_cpu, nthreads=cores, rep_offset=rep_offset, **kw)
    dt = time.perf_counter() - t0
    return n_cpu / dt, c, n_cpu, cores


def run_reference(args, spec):
    """CPU arm: the oracle (port of the reference) with all host threads on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as O
    mdl, kw, desc = spec["model"], spec["kw"], spec["desc"]
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
            "config": {"workload": desc, "replicates_per_step": nrep, "k": kw["k"], "m": kw["m"], "lag": kw["lag"]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference's drivers are interpreted R (no R in this image): timed the C++ oracle restatement, "
                    "which reproduces R's own README output on the emulated R stream and is faster than the interpreted "
                    "loop it restates"}
    print(json.dumps(line), flush=True)


class Bench:
    """One process per GPU: engine, NCCL plumbing, timing helpers."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from unbiasedmcmc_b200.engine import Engine
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.eng = Engine(self.local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def timed(self, spec, nrep, steps, warmup, step0=0, rep_offset_fn=None, check_nccl=False):
        """`warmup` untimed + `steps` timed plan launches of `nrep` replicates on this rank; -> dict of measurements.
        Timed region: barrier + sync | ev0 | K x (launch, fetch summary [D2H], NCCL all-reduce on the plan's stream) | ev1
        | sync + barrier; device time = max over ranks."""
        torch, dist = self.torch, self.dist
        mdl, kw = spec["model"], spec["kw"]
        plan = self.eng.plan(mdl, max_nrep=nrep, **kw)
        stream = torch.cuda.ExternalStream(plan.stream(), device=self.dev)
        ptr, slen = plan.summary_dev()

        class _Holder:   # zero-copy torch view of the plan's summary buffer
            __cuda_array_interface__ = {"shape": (slen,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        summary_t = torch.as_tensor(_Holder(), device=self.dev)
        reduced = torch.zeros(slen, dtype=torch.float64, device=self.dev)
        off = rep_offset_fn or (lambda step: (step * self.world + self.rank) * nrep)
        for w in range(warmup):
            plan.launch(seed=W.SEED, nrep=nrep, rep_offset=off(step0 + w))
        self.barrier()
        stream.synchronize()
        launches0 = self.eng.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kernel_ms, ops, cost, local_sum = 0.0, 0.0, 0.0, np.zeros(slen)
        self.barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for s in range(steps):
            plan.launch(seed=W.SEED, nrep=nrep, rep_offset=off(step0 + warmup + s))
            if self.world > 1:
                # the path's only collective: a few KB of sums over NVLink, ordered after the batch on the plan's own
                # stream and reduced into a separate buffer (the next launch rewrites the plan's summary)
                with torch.cuda.stream(stream):
                    reduced.copy_(summary_t)
                    dist.all_reduce(reduced)
            S = plan.fetch()["summary"]      # waits for the stream; the step's result to the host (D2H)
            km, _ = plan.last_ms()
            kernel_ms += km
            ops += spec["flops"](S)
            cost += S[4]
            local_sum = S
        ev1.record(stream)
        stream.synchronize()
        self.barrier()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        launches = self.eng.launch_count() - launches0
        dev_ms_max, wall_ms_max = self.max_over_ranks(dev_ms, wall * 1e3)
        out = {"dev_ms": dev_ms, "dev_ms_max": dev_ms_max, "wall_ms_max": wall_ms_max, "kernel_ms": kernel_ms, "ops": ops,
               "cost": cost, "launches": launches, "summary": local_sum}
        if check_nccl and self.world > 1:
            # the reduced summary of the LAST step against the sum of the per-rank host copies
            gathered = [torch.zeros(slen, dtype=torch.float64, device=self.dev) for _ in range(self.world)]
            dist.all_gather(gathered, torch.from_numpy(np.ascontiguousarray(local_sum)).to(self.dev))
            total = torch.stack(gathered).sum(0).cpu().numpy()
            red = reduced.cpu().numpy()
            n_all = float(red[0] + red[1])
            # counts, integer-valued sums (tau, cost, iteration) exactly; floating sums to the order of the reduction
            exact = bool(np.array_equal(red[:2], total[:2]) and np.array_equal(red[[2, 4, 5]], total[[2, 4, 5]]))
            close = bool(np.allclose(red, total, rtol=1e-12, atol=1e-9))
            out["nccl_check"] = {"ok": bool(exact and close and n_all == self.world * nrep), "n_replicates_reduced": n_all,
                                 "expected": self.world * nrep, "integer_sums_exact": exact, "float_sums_rtol_1e-12": close,
                                 "summary_len": int(slen)}
        plan.close()
        return out

    def e2e(self, spec, nrep, steps, step0):
        """the public host-buffer call, wall clock, same batch size; returns (value, last result, h2d, d2h)"""
        mdl, kw = spec["model"], spec["kw"]
        res = None
        # one untimed call of a small batch first: the first host-buffer call of a process pays one-off costs (context-level
        # allocator growth, first touch of the pinned staging the driver keeps for pageable copies) that are not per step
        self.eng.sample_unbiasedestimator(mdl, R.Draws(seed=W.SEED + 1), nrep=nrep, want_bins=False, max_iterations=kw["lag"] + 16,
                                          **{**kw, "k": kw["lag"] + 1, "m": kw["lag"] + 8})     # same buffers, 16 capped iterations
        self.barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            res = self.eng.sample_unbiasedestimator(mdl, R.Draws(seed=W.SEED), nrep=nrep,
                                                    rep_offset=((step0 + s) * self.world + self.rank) * nrep,
                                                    want_bins=False, **kw)
        self.barrier()
        wall, = self.max_over_ranks(time.perf_counter() - t0)
        dimh = mdl.h_dim(kw["h_kind"])
        nbins = kw["bins"].nbins if kw["bins"] is not None else 0
        # model parameters uploaded by every host-buffer call (what the ubm_model_* struct points at)
        h2d = sum(a.nbytes for a in getattr(mdl, "_flat", [])) if hasattr(mdl, "_flat") else 0
        h2d += sum(v.nbytes for k_, v in vars(mdl).items() if isinstance(v, np.ndarray) and not k_.startswith("chol"))
        if kw["bins"] is not None:
            h2d += kw["bins"].breaks.nbytes
        d2h = nrep * (3 * dimh + 3) * 8 + A.summary_len(dimh, nbins) * 8
        return self.world * nrep * steps / wall, res, int(h2d), int(d2h)


def roofline_of(name, spec, meas, peaks):
    bound = spec["bound"]
    achieved = meas["ops"] / (meas["kernel_ms"] * 1e-3) * 1e-12
    peak = peaks["fp64_tflops"] if bound == "fp64" else peaks["int32_tops"]
    r = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s" if bound == "fp64" else "Top/s",
         "frac": achieved / peak, "traffic": None, "kernel": W.KERNEL_OF[name],
         "kernel_share_of_step": meas["kernel_ms"] / meas["dev_ms"]}
    if name == "c2":
        r["kernel"] = "mvn_driver_kernel + mvn_single_kernel"
        r["kernel_note"] = ("two driver launches per step: the coupled pass hands every replicate over at its meeting time, the single-chain "
                            "continuation (32 slots per CTA) runs it to its horizon and takes about 80 % of the time; achieved = the step's "
                            "algorithmic flops over both launches")
    for fn in ("r02_traffic.json", "r01_traffic.json"):
        tj = os.path.join(ROOT, "profiles", fn)
        if os.path.exists(tj):
            tr = json.load(open(tj)).get(W.KERNEL_OF[name])
            if tr:   # DRAM bytes per launch, scaled from the committed ncu capture (bytes per replicate x replicates)
                r["traffic"] = tr["dram_bytes_per_replicate"] * meas["nrep"]
                r["traffic_source"] = tr["source"]
                break
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--nrep", type=int, default=0, help="replicates per GPU per step (default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --nrep (default 1e5 for c2) is the TOTAL per step, split over the ranks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-per-workload", action="store_true", help="headline workload only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    spec = W.bench_spec(args.workload)
    if args.impl == "reference":
        run_reference(args, spec)
        return

    B = Bench()
    rank, world = B.rank, B.world
    nrep = args.nrep or spec["nrep"]
    off_fn = None
    if args.scaling == "strong":
        total = nrep
        o, c = split_replicates(total, world)[rank]
        nrep = c
        off_fn = lambda step, o=o, total=total: step * total + o
    peaks = B.eng.measure_peaks()
    sampler = ClockSampler(B.local_rank)
    if rank == 0:
        sampler.start()
    meas = B.timed(spec, nrep, args.steps, args.warmup, rep_offset_fn=off_fn, check_nccl=True)
    clocks = sampler.stop() if rank == 0 else None
    meas["nrep"] = nrep
    nrep_all = world * nrep if args.scaling == "weak" else sum(c for _, c in split_replicates(args.nrep or spec["nrep"], world))
    value = nrep_all * args.steps / (meas["dev_ms_max"] * 1e-3)

    e2e_steps = max(1, min(args.steps, 2))
    e2e_value, res, h2d, d2h = B.e2e(spec, nrep, e2e_steps, args.warmup)

    # ---- all five BASELINE configs at the reference's horizons -----------------------------------------------------------
    per = {}
    if not args.no_per_workload and args.scaling == "weak":
        for name in ("c1", "c3", "c4", "c5"):
            if name == args.workload:
                continue
            sp = W.bench_spec(name)
            m_ = B.timed(sp, sp["nrep"], 1, 1)
            m_["nrep"] = sp["nrep"]
            ent = {"workload": sp["desc"], "value": world * sp["nrep"] / (m_["dev_ms_max"] * 1e-3), "unit": UNIT,
                   "replicates_per_gpu_per_step": sp["nrep"], "k": sp["kw"]["k"], "m": sp["kw"]["m"], "lag": sp["kw"]["lag"],
                   "steps": 1, "warmup": 1, "ms_per_step": m_["dev_ms_max"], "gpu_launches": int(m_["launches"]),
                   "mean_meetingtime": float(m_["summary"][2] / max(m_["summary"][0], 1.0)),
                   "kernel_steps_per_s_rank0": m_["cost"] / (m_["dev_ms"] * 1e-3),
                   "roofline": roofline_of(name, sp, m_, peaks)}
            if world == 1:
                e2e_n = 1 if name == "c4" else 2          # a C4 step is 15 s; two steps for the others (host-side jitter)
                ev, r_, h2d_, d2h_ = B.e2e(sp, sp["nrep"], e2e_n, 2)
                ent["e2e"] = {"value": ev, "unit": UNIT, "h2d_bytes_per_step": h2d_, "d2h_bytes_per_step": d2h_, "steps": e2e_n}
                if not args.no_cpu:
                    cv, c, n_cpu, cores = cpu_sample(sp, 6.0, rep_offset=(2 + e2e_n - 1) * sp["nrep"])     # the last e2e step's replicates
                    ent["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
                                           "sample": f"first {n_cpu} replicates of the last e2e step, oracle C++ port, {cores} std::threads",
                                           "bit_identical_to_gpu": bool(np.array_equal(c["uestimator"], r_["uestimator"][:n_cpu]))}
            per[name] = ent
        if world > 1 and args.workload == "c2":
            # strong scaling of the headline config: BASELINE's nrep = 1e5 in TOTAL, so the last-wave tail of the heavy-tailed
            # meeting times shows
            total = 100000
            o, c = split_replicates(total, world)[rank]
            m_ = B.timed(spec, c, 1, 1, rep_offset_fn=lambda step, o=o: step * total + o)
            per["c2_strong"] = {"workload": spec["desc"], "scaling": "strong", "replicates_total_per_step": total,
                                "value": total / (m_["dev_ms_max"] * 1e-3), "unit": UNIT, "ms_per_step": m_["dev_ms_max"],
                                "steps": 1, "warmup": 1}

    if rank != 0:
        if world > 1:
            B.dist.destroy_process_group()
        return

    roofline = roofline_of(args.workload, spec, meas, peaks)
    roofline["peak_source"] = ("measured live by ubm_measure_peaks (independent DFMA / IMAD chains, 2 ops each); "
                               "MEASURED_PEAKS.json holds no FP64 / INT32 figure; HBM is not the bound (state stays on chip, "
                               "dram traffic of the driver kernel is a few KB per launch, see profiles/)")
    roofline["fp64_tflops_measured"] = peaks["fp64_tflops"]
    roofline["int32_tops_measured"] = peaks["int32_tops"]
    roofline["fp64_tensor_tflops_measured"] = peaks["fp64_tensor_tflops"]   # mma.m8n8k4.f64 chains: the same pipe, the same peak
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        roofline["hbm_gbs_measured_peak"] = json.load(open(mp)).get("hbm_gbs")

    cpu = None
    if world == 1 and not args.no_cpu:
        cv, c, n_cpu, cores = cpu_sample(spec, 15.0, rep_offset=args.warmup * nrep, nrep_cap=nrep)
        g = B.eng.sample_unbiasedestimator(spec["model"], R.Draws(seed=W.SEED), nrep=n_cpu, rep_offset=args.warmup * nrep,
                                           want_bins=False, **spec["kw"])
        cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {n_cpu} replicates of the first timed step, oracle C++ port, {cores} std::threads",
               "bit_identical_to_gpu": bool(np.array_equal(c["uestimator"], g["uestimator"]))}

    kw = spec["kw"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": meas["dev_ms_max"] / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": spec["desc"], "replicates_per_gpu_per_step": nrep, "k": kw["k"], "m": kw["m"], "lag": kw["lag"],
                       "l2": "no reuse across steps: chain state lives in registers/shared memory; each step draws "
                             "fresh Philox streams, outputs (nrep x dimh doubles) are written once",
                       "timing": "CUDA events on the engine's stream (kernels, D2H of the summary and, for N > 1, the NCCL "
                                 "all-reduce enqueued on that stream), max over ranks; wall_ms_max also given"},
            "wall_ms_per_step": meas["wall_ms_max"] / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "call": "ubm_sample_unbiasedestimator (host buffers, synchronous)"},
            "gpu_launches": int(meas["launches"]), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "kernel_steps_per_s_rank0": meas["cost"] / (meas["dev_ms"] * 1e-3),
            "mean_meetingtime": float(meas["summary"][2] / max(meas["summary"][0], 1.0))}
    if "nccl_check" in meas:
        line["nccl_check"] = meas["nccl_check"]
    if per:
        line["per_workload"] = per
    print(json.dumps(line), flush=True)
    if world > 1:
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
