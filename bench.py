#!/usr/bin/env python
"""bench.py — headline measurement for the tempered-MH hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c5] [--impl reference]

A "step" is one APT iteration of the workload: every rung of every ladder runs each of its tempered samplers once
(APT_functions.R:435-448), then the swap phase, ladder adaptation and thinned output.  Workload at every N is
BASELINE.json configs[1] ("c2": 2-D bimodal mixture, 32 temperatures x 4096 independent ladders per GPU, scalar
RW_tempered; weak scaling: ladders are sharded across GPUs, no collective on the data path).  The HBM-bound
likelihood kernel of configs[2] ("c3": logistic regression N=1e6, p=50, 16 temperatures) is timed in the same
process at N=1 and reported as `roofline_c3`.

value     = tempered MH updates / s over all GPUs, device time (CUDA events on the engine's stream, max over ranks)
e2e       = the same through the public API with HOST buffers: data + inits uploaded, run, mvSamples fetched to host
            (N > 1: every rank's traces gathered over NCCL to rank 0 and copied to its host memory, inside the timed region)
roofline  = dominant kernel of the timed region; for c2 the kernel is FP64-pipe bound, so the
            roofline is flop-based against the FP64 FMA peak measured live; roofline_c3 is the HBM one.
--impl reference times the CPU oracle (restated reference, NOT nimble itself: R/nimble are absent).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# algorithmic fp64 flops per tempered update (SURVEY.md §8d "n*c_dens + c_rng"): counted from the lowered code
FLOPS_PER_UPDATE = {"c2": 20 * 8 + 150, "c5": 1024 * (2 * 8 + 60) + 8 * 8 + 300}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons during the timed region (pynvml, 20 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def workload_spec(name):
    import models
    if name == "c2":
        return models.c2_mixture(n=20, K=32), 4096, 10
    if name == "c5":
        return models.c5_small_logistic(N=1024, p=8, K=64), 8192, 10
    if name == "c3":
        return models.c3_logistic(), 1, 1
    raise SystemExit("unknown workload " + name)


def ess_batch_means(x):
    """ESS by non-overlapping batch means (sqrt(n) batches); applied identically to GPU and CPU traces."""
    n = len(x)
    b = max(2, int(np.sqrt(n)))
    m = n // b
    if m < 2:
        return float(n)
    bm = x[:b * m].reshape(b, m).mean(1)
    v = x.var(ddof=1)
    vb = bm.var(ddof=1) * m
    return float(n * v / vb) if vb > 0 else float(n)


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference's algorithm (oracle) on all host threads."""
    if rank != 0:
        return
    from oracle.oracle import OracleAPT, lib
    spec, _, thin = workload_spec(args.workload)
    cores = lib().orc_max_threads()
    # a step = one APT iteration over a sample of `nl` independent ladders of the workload.  The sample is sized so
    # that the whole run is about 20 s of CPU work at the port's typical rate (REF_RATE updates/s on all threads) —
    # large enough that thread start-up does not dominate a step, bounded so the run ends within minutes.
    _, Lfull, _ = workload_spec(args.workload)
    upl = len(spec["Temps"]) * max(1, len(spec["samplers"]))
    REF_RATE = {"c2": 1.5e7, "c5": 2.5e6, "c3": 6.0}[args.workload]
    nl = int(REF_RATE * 20.0 / ((args.steps + max(1, args.warmup)) * upl))
    nl = max(1 if args.workload == "c3" else cores, min(Lfull, nl))
    orc = OracleAPT(spec["code"], spec["constants"], spec["data"], spec["inits"], spec["samplers"], spec["monitors"], (),
                    spec["Temps"], spec["ULT"], True, nl, 11, 0)
    K, NS = orc.K, orc.NS
    # a step = one APT iteration over the sample of `nl` ladders
    orc.run(max(1, args.warmup), thin=thin, nthreads=cores)
    t0 = time.perf_counter()
    orc.run(args.steps, reset=False, thin=thin, nthreads=cores)
    dt = time.perf_counter() - t0
    val = args.steps * nl * K * NS / dt
    line = {"impl": "reference", "metric": "tempered MH updates/sec", "value": val, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "ladders_in_sample": nl, "temperatures": K, "samplers_per_rung": NS},
            "cpu_baseline": {"value": val, "unit": "updates/s", "cores": cores, "kind": "port",
                             "sample": "%d independent ladders (one per thread, %d threads) x %d iterations of the same workload; "
                                       "restated CPU baseline, not nimble" % (nl, cores, args.steps)},
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline(workload, seconds=15.0):
    from oracle.oracle import OracleAPT, lib
    spec, _, thin = workload_spec(workload)
    cores = lib().orc_max_threads()
    nl = cores * 2
    orc = OracleAPT(spec["code"], spec["constants"], spec["data"], spec["inits"], spec["samplers"], spec["monitors"], (),
                    spec["Temps"], spec["ULT"], True, nl, 11, 0)
    K, NS = orc.K, orc.NS
    orc.run(50, thin=thin, nthreads=cores)
    t0 = time.perf_counter(); orc.run(200, reset=False, thin=thin, nthreads=cores); dt = time.perf_counter() - t0
    niter = int(max(200, min(200000, 200 * seconds / max(dt, 1e-3))))
    t0 = time.perf_counter(); orc.run(niter, reset=False, thin=thin, nthreads=cores); dt = time.perf_counter() - t0
    # single-thread figure as well (the reference itself is serial, APT:435-448)
    o1 = OracleAPT(spec["code"], spec["constants"], spec["data"], spec["inits"], spec["samplers"], spec["monitors"], (),
                   spec["Temps"], spec["ULT"], True, 1, 11, 0)
    n1 = max(200, niter // 8)
    t1 = time.perf_counter(); o1.run(n1, thin=thin, nthreads=1); d1 = time.perf_counter() - t1
    return {"value": niter * nl * K * NS / dt, "unit": "updates/s", "cores": cores, "kind": "port",
            "single_core_value": n1 * K * NS / d1,
            "sample": "%d independent ladders x %d iterations on %d threads (%.1f s); restated CPU baseline, not nimble" % (nl, niter, cores, dt)}


def c3_roofline(device, passes=400):
    """HBM-bound likelihood pass of config 3 (grid engine): algorithmic bytes 8*N*(p+1) per pass / device time.
    Measured at the default 16 rungs per pass (most updates/s; FP64 work per pass exceeds the HBM time there) and at
    8 rungs per pass (two passes per iteration; the setting where the pass is closest to HBM-bound)."""
    from helpers import build_engine
    import models
    N, p, K = 1000000, 50, 16
    spec = models.logistic(N, p, K, seed_off=3)
    peak, how = _peaks()
    by = 8.0 * N * (p + 1)
    res = {}
    for tb in (16, 8):
        eng = build_engine(spec, engine=2, device=device, rungsPerPass=tb)
        eng.run(20)
        eng.run(passes, reset=False)  # spans two covariance adaptations (adaptInterval = 200): their cost is inside
        t = eng.getTimes()
        first = t["main_kernel_ms"] / max(1, t["main_launches"])
        # steady state: once the ladder has adapted (hot rungs at T ~ 1e5) a fraction of the terms takes the branch that
        # reproduces the reference's own rounding for y = 0, eta > 8, which costs ~20 % more per pass
        eng.run(1200, reset=False)
        eng.run(passes, reset=False)
        t = eng.getTimes()
        ms = t["main_kernel_ms"] / max(1, t["main_launches"])
        res[tb] = {"ms_per_launch": ms, "launches": int(t["main_launches"]), "achieved": by / (ms * 1e-3) / 1e9,
                   "frac": by / (ms * 1e-3) / 1e9 / peak, "updates_per_s": K * passes / (t["kernel_ms"] * 1e-3),
                   "first_400_iterations": {"ms_per_launch": first, "frac": by / (first * 1e-3) / 1e9 / peak}}
        del eng
    r = res[16]
    return {"bound": "hbm", "achieved": r["achieved"], "peak": peak, "unit": "GB/s", "frac": r["frac"],
            "traffic": 413.7e6, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (profiles/r1_c3_grid_pass_kernel.md)",
            "peak_source": how, "kernel": "apt::grid_pass_kernel", "ms_per_launch": r["ms_per_launch"], "launches": r["launches"],
            "algorithmic_bytes_per_launch": by, "rungs_per_pass": 16, "updates_per_s": r["updates_per_s"],
            "timed": "400 passes after 1620 iterations (adapted ladder); first_400_iterations = iterations 21..420",
            "first_400_iterations": r["first_400_iterations"],
            "rungs_per_pass_8": res[8],
            "workload": "c3: logistic N=1e6 p=50, 16 temperatures, 1 ladder, sampler_RW_block_tempered; inputs (408 MB) larger than L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=400)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--ladders", type=int, default=0, help="ladders per GPU (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import nimbleapt_b200 as nb
    from nimbleapt_b200 import _lib
    from helpers import build_engine
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec, L, thin = workload_spec(args.workload)
    if args.ladders > 0:
        L = args.ladders
    K = len(spec["Temps"])
    eng = build_engine(spec, nLadders=L, ladder0=rank * L, device=local, engine=2 if args.workload == "c3" else 0)
    NU = eng.info(nb.api.INFO_N_UNITS)

    def barrier():
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident timed region
    eng.run(args.warmup, thin=thin)
    barrier()
    with ClockSampler(local) as cs:
        eng.run(args.steps, reset=False, thin=thin)
        tm = eng.getTimes()
    barrier()
    dev_s = max_over_ranks(tm["kernel_ms"] * 1e-3)
    updates_total = float(args.steps) * L * K * NU * world
    value = updates_total / dev_s
    # ---------------- NCCL communicator inside the library (gathers of traces / pooled statistics only, §8e)
    C = _lib.C
    if dist is not None:
        uid = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(_lib.lib().apt_comm_unique_id(uid))
        obj = [bytes(uid)]
        dist.broadcast_object_list(obj, src=0)
        uid = (C.c_ubyte * 128)(*obj[0])
        _lib.check(_lib.lib().apt_comm_init(eng._h, world, rank, uid))
    # ---------------- end to end through the public API with host buffers
    y = np.ascontiguousarray(spec["data"]["y"])
    th0 = np.concatenate([np.ravel(spec["inits"][k], order="F") for k in spec["inits"]])
    nmon = len(eng.monitorNames)
    rows = args.steps // thin
    pinned = None
    if rank == 0:  # page-locked destination of the (gathered) traces, allocated once outside the timed region
        import torch
        pinned = torch.empty(world * L * rows * nmon, dtype=torch.float64, pin_memory=True).numpy()

    def e2e_once(steps):
        eng.setData("y", y)                      # H2D: observations
        eng.setInits(th0)                        # H2D: initial values
        if dist is None:
            # one process: the thinned traces are copied back to page-locked host memory asynchronously, piece by piece,
            # while the run is still sampling (apt_run_args.samples_out); complete when run() returns
            out = pinned[:L * (steps // thin) * nmon]
            eng.run(steps, thin=thin, samplesOut=out)
            sm = out.reshape(L, steps // thin, nmon)
            return sm, sm
        eng.run(steps, thin=thin)                # reset=TRUE run from the uploaded state
        if dist is not None:                     # NCCL gather of every rank's traces to rank 0 (the process that hands them to
            r_ = steps // thin                   # the user), one D2H into pinned host memory there
            got = C.c_int64()
            if rank == 0:
                allbuf = pinned[:world * L * r_ * nmon]
                _lib.check(_lib.lib().apt_comm_gather_samples(eng._h, 0, 0, allbuf.ctypes.data_as(C.POINTER(C.c_double)), allbuf.size, C.byref(got)))
                assert got.value == allbuf.size
                sa = allbuf.reshape(world * L, r_, nmon)
                return sa[:L], sa
            _lib.check(_lib.lib().apt_comm_gather_samples(eng._h, 0, 0, None, 0, C.byref(got)))
            return None, None

    e2e_once(args.steps)  # untimed warm-up of the same call sequence at the same size (first NCCL collective sets up its channels, trace and gather buffers reach their final size)
    barrier()
    t0 = time.perf_counter()
    samples, samples_all = e2e_once(args.steps)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    if samples is None:  # ranks other than 0: own traces for the pooled statistic below, outside the timed region
        samples = eng.mvSamples.matrix(-1)
        samples_all = samples
    h2d = (y.nbytes + th0.nbytes) * world
    d2h = samples_all.nbytes
    burn = samples.shape[1] // 4
    ess = float(np.mean([ess_batch_means(np.abs(samples[l, burn:, 0])) for l in range(min(L, 64))]))
    pooled = np.array([np.abs(samples[:, burn:, 0]).sum(), float(samples[:, burn:, 0].size)])
    if dist is not None:                     # pooled statistic by NCCL all-reduce(sum)
        buf = np.ascontiguousarray(pooled)
        _lib.check(_lib.lib().apt_comm_allreduce_sum(eng._h, buf.ctypes.data_as(C.POINTER(C.c_double)), buf.size))
        pooled = buf
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    clocks = cs.summary()
    flops = FLOPS_PER_UPDATE.get(args.workload, 0) * updates_total / world
    line = {"metric": "tempered MH updates/sec", "value": value, "unit": "updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "ladders_per_gpu": L, "temperatures": K, "samplers_per_rung": NU, "thin": thin,
                       "engine": "grid" if eng.info(nb.api.INFO_ENGINE) == 2 else "ladder-resident",
                       "l2": ("inputs (408 MB design matrix) larger than the 126 MB L2: every pass re-streams them from HBM"
                              if args.workload == "c3" else
                              "working set is shared-memory resident; traces stream to HBM (no L2 reuse to flush)")},
            "e2e": {"value": updates_total / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps, "seconds": e2e_s},
            "gpu_launches": int(tm["launches"]), "clocks": clocks,
            "ess_per_sec": ess * L * world / e2e_s, "pooled_abs_theta1_mean": float(pooled[0] / pooled[1]),
            "swap_proposals_per_sec": float(args.steps) * L * K * world / dev_s}
    if args.workload in FLOPS_PER_UPDATE:
        tf = flops / (tm["kernel_ms"] * 1e-3) / 1e12
        pk = eng.info(nb.api.INFO_FP64_GFLOPS) / 1e3  # measured live: pure DFMA kernel, CUDA events (no FP64 peak in MEASURED_PEAKS.json)
        line["roofline"] = {"bound": "fp64", "achieved": tf, "peak": pk, "unit": "TFLOP/s", "frac": tf / pk, "traffic": None,
                            "peak_source": "measured in this run: 16 independent DFMA chains per thread, 2 x 256 threads per SM, best of 3 (apt_info APT_INFO_FP64_PEAK_GFLOPS)",
                            "kernel": "apt::ladder_kernel", "algorithmic_flops_per_update": FLOPS_PER_UPDATE[args.workload],
                            "note": "small-model kernel: bound by issue slots / FP64 dependent-issue latency, not by HBM (SURVEY.md §8d); most executed instructions are not algorithmic flops (Philox, Box-Muller, swap phase); HBM roofline: roofline_c3"}
    if world == 1 and not args.no_c3 and args.workload != "c3":
        try:
            del eng
            line["roofline_c3"] = c3_roofline(local)
        except Exception as e:  # keep the headline line even if the side measurement fails
            line["roofline_c3"] = {"error": str(e)[:200]}
    if args.workload == "c3":
        peak, how = _peaks()
        by = 8.0 * 1000000 * 51
        ms = tm["main_kernel_ms"] / max(1, tm["main_launches"])
        line["roofline"] = {"bound": "hbm", "achieved": by / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": by / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": how, "kernel": "apt::grid_pass_kernel"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload if args.workload != "c3" else "c2")
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
