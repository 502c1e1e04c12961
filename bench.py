#!/usr/bin/env python3
"""bench.py -- TikTak hot path on B200: Sobol objective evals/s and local restarts/s.

    python bench.py --gpus N --steps K --warmup W [--workload C2|C3|C4|C5] [--impl reference] [--no-extra]

One "step" = one pass of the hot path over one synthetic job: Sobol pre-test of qr_ndraw points (solveAtSobolPoints) ->
start selection (chooseSobol) -> all local restarts (LocalMinimizations: rounds of round_size restarts; later rounds are
started early and kept while the best row is unchanged -- same rows as strictly sequential rounds), device-resident.

Headline line (the driver's contract): N = 1 runs BASELINE.json configs[1] (C2: Griewank 10-D, 1e6 Sobol points, 1001
Nelder-Mead restarts).  N > 1 (torchrun, one rank per GPU; the library's own NCCL communicator does the exchanges) shards the
Sobol index range and every launch's restarts over the ranks; per-GPU work is held fixed: "scaling": "weak".

Beyond the base contract the line carries
  restarts_per_s    local restarts/s of the same step (second half of BASELINE.json's metric), roofline_local for its kernel
  roofline          the Sobol evaluation kernel: algorithmic FP64 flops per evaluation x evaluations per launch /
                    CUDA-event duration on the library's stream, against the FP64 DFMA peak measured on this GPU
  cpu_baseline      the oracle port timed on this box's host cores on a bounded sample
  e2e               the same metric through the public host API: the stage's counters read back every step, all stored values
                    streamed to pinned host memory (second stream, beside selection and restarts)
  per_workload      N = 1: BASELINE.json configs[2..4] (C3, C4, C5 at BASELINE.md's sizes), each with its Sobol-kernel time and
                    roofline fraction, restarts/s and a small cpu_baseline -- one or two steps each
  strong_c3         configs[2] EXACTLY as stated (1e8 points, 10 001 restarts) over the N GPUs of this run: whole-step time;
                    N = 1, 2, 4, 8 of this block are the strong-scaling curve
  parity            a small sharded job of this very run (N ranks) compared with the single-instance CPU oracle on rank 0,
                    outside every timed region: driver-visible multi-GPU parity
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic FP64 flops per Sobol evaluation, SURVEY.md section 8(d) / BASELINE.md section 3
F_ALG = {"C1": 237, "C2": 545, "C3": 1105, "C4": 905, "X7": 383}  # X7: Griewank 7-D by the same per-coordinate count (54 nx + 5)
OBJ_NAME = {"C1": "template", "C2": "Griewank", "C3": "Rastrigin", "C4": "Rosenbrock", "C5": "synthetic SMM", "X7": "Griewank"}
# restarts per round (all blended against the best row of the round's start)
ROUND_SIZE = {"C1": 4, "C2": 128, "C3": 1024, "C4": 3552, "C5": 128}  # C4: the 24 warps per SM the one-warp kernel keeps resident (round 1: 1184)
# asynchronous instances (restart stage, one GPU): how many blocks / warps play the reference's processes -- about two restarts per
# instance: 512 four-warp blocks at C2 (1.3 ms; 256: 1.9 ms, the same best value), every resident WARP at C3 (4096) and C4 (3548) -- asking
# for at least 12 per SM selects the one-warp instances (tiktak_cuda_local_async_clock) -- and C5 (147: one block per SM, one SM keeps the
# time); the library clamps to what is resident
ASYNC_INSTANCES = {"C1": 4, "C2": 512, "C3": 4096, "C4": 4096, "C5": 147}
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the stage's kernel, from the committed `ncu --set full` captures
TRAFFIC = {"C2": (25088, "profiles/r01d_full_c2_sobol.csv"), "C3": (742264320, "profiles/r01d_full_c3_sobol.csv"),
           "C4": (21494528, "profiles/r01d_full_c4_sobol.csv"), "C5": (1020513512, "profiles/r01c_full_c5_smm.csv (1184 evaluations)")}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--round-size", type=int, default=0)
    ap.add_argument("--local-mode", choices=["auto", "rounds", "async"], default="auto",
                    help="restart stage: rounds of round_size restarts, or round_size asynchronous instances (one launch; the reference's "
                         "multi-process mode); auto = async where the library has it (Nelder-Mead, nx <= 31, one GPU)")
    ap.add_argument("--nm-shrink-mode", type=int, default=-1, choices=[-1, 0, 1],
                    help="initial-simplex rule: 0 = the reference code, 1 = README.md:36 (bounds from the best minima so far); "
                         "default 0: the rule is not in the reference's code (SURVEY finding 3)")
    ap.add_argument("--no-local", action="store_true", help="time only the Sobol stage + selection")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only: no per_workload / strong_c3 / parity blocks")
    return ap.parse_args()


def workload_config(name, world, round_size=0, nm_shrink_mode=-1, weak=True):
    from tiktak_b200 import baseline_config

    cfg = baseline_config(name)  # C5: BASELINE.md section 2 (panel 2^17 x 40, 200 000 draws, 100 000 legitimate, 1000 + 1 restarts)
    cfg.nm_shrink_mode = 0 if nm_shrink_mode < 0 else nm_shrink_mode
    scale = world if weak else 1
    if scale > 1:  # weak scaling: per-GPU work fixed
        cfg.qr_ndraw *= scale
        cfg.maxpoints *= scale
        cfg.legitimate = cfg.legitimate * scale if name == "C5" else cfg.qr_ndraw + 10000
        cfg.lottery_points = cfg.maxpoints
    cfg.round_size = (round_size or ROUND_SIZE[name]) * scale
    return cfg


def workload_string(name, cfg):
    alg = "DFNLS" if name == "C5" else "Nelder-Mead"
    return (f"{name}: {OBJ_NAME[name]} {cfg.nx}-D, nm={cfg.nm}, qr_ndraw={cfg.qr_ndraw}, legitimate={cfg.legitimate}, "
            f"maxpoints={cfg.maxpoints} (+1 = {cfg.p_maxpoints} {alg} restarts), round_size={cfg.round_size}, "
            f"nm_itmax={cfg.nm_itmax}, ftol={cfg.nm_ftol}, nm_shrink_mode={cfg.nm_shrink_mode}")


# ------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons through NVML while the timed region runs"""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.active = False  # samples are kept only while a timed region runs
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                        self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(
                        self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.004)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    name, npts, nrestarts = args
    sys.path.insert(0, ROOT)
    from oracle.oracle import MATH_LIBM, Oracle

    cfg = workload_config(name, 1)
    cfg.qr_ndraw = npts
    cfg.maxpoints = min(nrestarts, npts)
    cfg.legitimate = npts + 10000
    o = Oracle(cfg, MATH_LIBM)  # libm cos: what the reference's intrinsic COS is
    t0 = time.perf_counter()
    o.solveAtSobolPoints()
    t1 = time.perf_counter()
    if name == "C5":  # one DFNLS restart costs 320 panel simulations (~1 min/core): not sampled, see `sample`
        return t1 - t0, 0.0, npts, 0, 320.0
    o.chooseSobol(mode=0)
    t2 = time.perf_counter()
    o.LocalMinimizations(1, round_size=1)
    t3 = time.perf_counter()
    return t1 - t0, t3 - t2, npts, cfg.maxpoints + 1, float(np.mean(o.evals))


def _cpu_worker_faithful(args):
    """Sobol stage with the reference's per-point file protocol (oracle: orc_pretest_faithful) in a private directory"""
    import ctypes as C
    import tempfile

    name, npts = args
    sys.path.insert(0, ROOT)
    from oracle.oracle import MATH_LIBM, Oracle

    cfg = workload_config(name, 1)
    cfg.legitimate = cfg.qr_ndraw + 10000
    o = Oracle(cfg, MATH_LIBM)
    o.L.orc_pretest_faithful.restype = C.c_long
    o.L.orc_pretest_faithful.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        n = o.L.orc_pretest_faithful(o.h, d.encode(), npts)
        t1 = time.perf_counter()
    return t1 - t0, int(n)


_CAL = {}


def cpu_calibrate(name):
    """seconds per Sobol evaluation and per restart of one oracle instance on this host (tiny sample)"""
    if name not in _CAL:
        n_cal = 3 if name == "C5" else 2000
        t_s, t_l, n0, r0, _ = _cpu_worker((name, n_cal, 2))
        _CAL[name] = (t_s / n0, (t_l / r0 if r0 else 1.0), n_cal)
    return _CAL[name]


def cpu_sample(name, npts, nres, faithful=False):
    """one bounded sample: `cores` independent oracle instances (the reference's own parallel model, README.md:30-33),
    compute-only (no .dat I/O), each npts Sobol points + nres + 1 restarts"""
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        outs = pool.map(_cpu_worker, [(name, npts, nres)] * cores)
        wall = time.perf_counter() - t0
        outs_f = pool.map(_cpu_worker_faithful, [(name, 20000)] * cores) if (faithful and name != "C5") else None
    ts, tl = max(o[0] for o in outs), max(o[1] for o in outs)
    evals_s = cores * npts / ts
    restarts_s = (cores * (nres + 1) / tl) if tl > 0 else evals_s / 320.0
    out = {"value": evals_s, "unit": "evals/s", "cores": cores, "kind": "port",
           "sample": f"{cores} independent oracle instances x ({npts} Sobol points of {name} + "
                     + (f"{nres + 1} Nelder-Mead restarts)" if name != "C5" else "no DFNLS restart: one costs 320 panel simulations, restarts/s = evals/s / 320)")
                     + f", libm cos, compute-only; wall {wall:.1f}s",
           "sample_ms": 1e3 * (ts + tl), "sample_wall_s": wall, "restarts_per_s": restarts_s,
           "mean_evals_per_restart": float(np.mean([o[4] for o in outs])),
           "single_core_evals_per_s": npts / float(np.mean([o[0] for o in outs]))}
    if outs_f:  # what the reference's users pay: ~6 file open/close cycles around every evaluation
        out["with_file_protocol"] = {"value": cores * 20000 / max(o[0] for o in outs_f), "unit": "evals/s",
                                     "sample": f"{cores} instances x 20000 Sobol points with the lastSobol/legitSobol/logFile/"
                                               f"sobolFnVal file cycles of TiktakGlobalSearch.f90:426-463 in a private temp directory"}
    return out


def cpu_arm(name, budget_s, faithful=True):
    """The reference cannot be compiled here (no Fortran compiler, ifort-only source); its CPU implementation of the path is
    timed as the line-faithful oracle port."""
    from oracle import oracle as orc

    orc.load()
    per_eval, per_restart, n_cal = cpu_calibrate(name)
    npts = int(max(n_cal, min(2_000_000, 0.5 * budget_s / per_eval)))
    nres = int(max(2, min(2000, 0.5 * budget_s / per_restart)))
    return cpu_sample(name, npts, nres, faithful)


def run_reference(args, rank, world):
    """the reference arm: the CPU implementation of the path (oracle port) on this box's host cores, on the GPU arm's config;
    every step is one bounded sample of that workload"""
    if rank != 0:
        return
    name = args.workload
    cfg = workload_config(name, world, args.round_size, args.nm_shrink_mode)
    per_eval, per_restart, n_cal = cpu_calibrate(name)
    total = max(1, args.steps + args.warmup)
    budget = max(1.0, min(args.cpu_seconds, 90.0 / total))  # the whole run ends within a couple of minutes
    npts = int(max(n_cal, min(2_000_000, 0.5 * budget / per_eval)))
    nres = int(max(2, min(2000, 0.5 * budget / per_restart)))
    for _ in range(args.warmup):
        cpu_sample(name, max(n_cal, npts // 8), max(2, nres // 8))
    steps = [cpu_sample(name, npts, nres) for _ in range(max(1, args.steps))]
    value = float(np.mean([s["value"] for s in steps]))
    cb = dict(steps[-1])
    cb["value"] = value
    cb["restarts_per_s"] = float(np.mean([s["restarts_per_s"] for s in steps]))
    line = {"impl": "reference", "metric": "sobol_objective_evals_per_s", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": len(steps), "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean([s["sample_wall_s"] for s in steps])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(name, cfg),
                       "parallelism": f"{cb['cores']} independent CPU instances (the reference's own parallel model)",
                       "note": "reference arm = CPU oracle port (the Fortran reference is not buildable in this image); each step is "
                               "a bounded sample of the workload: " + cb["sample"]},
            "restarts_per_s": cb["restarts_per_s"], "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- GPU arm
class Job:
    """one workload bound to this rank's GPU: timed steps of the hot path, stage times from CUDA events on the library stream"""

    def __init__(self, name, cfg, rank, world, local_rank, dist, no_local=False, local_mode="auto"):
        import torch

        from tiktak_b200 import SOBOL_BLOCK, TikTakSearch

        self.torch, self.dist, self.name, self.cfg, self.world, self.rank, self.no_local = torch, dist, name, cfg, world, rank, no_local
        self.dev = torch.device("cuda", local_rank)
        self.s = TikTakSearch(cfg, device=local_rank, rank=rank, world=world)
        self.ext = self.s.torch_stream()  # the library's stream: CUDA events recorded on it time the stages on the device
        self.n_host = ((cfg.qr_ndraw // SOBOL_BLOCK + world) // world) * SOBOL_BLOCK
        self.pinned = None  # host side of the e2e copy, allocated on first use
        self.alg = "b" if name == "C5" else "a"
        self.local_async = local_mode != "rounds" and world == 1 and self.s.asyncSupported(self.alg)
        if local_mode == "async" and not self.local_async:
            raise SystemExit("bench: --local-mode async is not available for this workload / world size")

    def close(self):
        self.s.close()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def mark(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.ext)
        return e

    def step(self, flush, e2e):
        """one pass; returns device stage times (ms, CUDA events on the library stream) and wall times (s)"""
        s = self.s
        flush.zero_()  # evict L2 between steps (untimed)
        self.barrier()
        t0 = time.perf_counter()
        _, ne, nl = s.solveAtSobolPoints(wait=e2e)  # e2e: the stage's counters (lastSobol / legitSobol) are read back
        d2h = 0
        if e2e:  # every stored value streams to pinned host memory on a second stream, beside selection and restarts
            if self.pinned is None:
                self.pinned = self.torch.empty(self.n_host, dtype=self.torch.float64, pin_memory=True).numpy()
            d2h = 8 * s.sobolFnValShardAsync(self.pinned) + 16
        t1 = time.perf_counter()
        e2 = self.mark()
        s.chooseSobol()
        e3 = self.mark()
        t3 = time.perf_counter()
        if not self.no_local:
            if self.local_async:
                try:
                    s.LocalMinimizationsAsync(self.alg, instances=ASYNC_INSTANCES[self.name])
                except Exception as exc:  # e.g. no cooperative launch on this device: the rounds form always exists
                    print(f"bench: asynchronous instances unavailable ({exc}); using rounds", file=sys.stderr)
                    self.local_async = False
                    s.LocalMinimizations(self.alg)
            else:
                s.LocalMinimizations(self.alg)
        e4 = self.mark()
        self.ext.synchronize()
        t4 = time.perf_counter()
        if e2e:
            s.sobolFnValWait()
        t5 = time.perf_counter()
        return dict(pretest_ms=s.stage_ms("pretest"), choose_ms=e2.elapsed_time(e3),
                    local_ms=0.0 if self.no_local else e3.elapsed_time(e4), local_kernels_ms=None if self.no_local else s.stage_ms("local"), pretest_eval_ms=s.stage_ms("pretest_eval"),
                    wall_pretest=t1 - t0, wall_choose=t3 - t1, wall_local=t4 - t3, wall=t5 - t0, copy_tail=t5 - t4,
                    ne=ne, nl=s.n_legit, d2h=d2h)

    def mx(self, v):  # max over ranks
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def measure(self, flush, steps, warmup, sampler=None, e2e_steps=0):
        for _ in range(warmup):
            self.step(flush, False)
        if e2e_steps:
            self.step(flush, True)  # first use of the pinned buffer and the copy stream: untimed
        if sampler is not None:
            sampler.active = True
        l0 = self.s.launch_count()
        dev = [self.step(flush, False) for _ in range(steps)]
        launches = (self.s.launch_count() - l0) // max(steps, 1)
        e2e = [self.step(flush, True) for _ in range(e2e_steps)]
        if sampler is not None:
            sampler.active = False
        s, cfg = self.s, self.cfg
        pre_ms = self.mx(np.mean([d["pretest_ms"] for d in dev]))
        choose_ms = self.mx(np.mean([d["choose_ms"] for d in dev]))
        local_ms = self.mx(np.mean([d["local_ms"] for d in dev]))
        eval_ms = self.mx(np.mean([d["pretest_eval_ms"] for d in dev]))
        n_eval = s.n_evaluated if s.n_evaluated else cfg.qr_ndraw  # the legitimate cut may stop early (C5)
        out = {"workload": workload_string(self.name, cfg), "n_evaluated": int(n_eval), "n_legit": int(s.n_legit), "steps": steps, "warmup": warmup,
               "stage_ms": {"pretest": pre_ms, "pretest_eval_kernel": eval_ms, "choose": choose_ms, "local": local_ms},
               "ms_per_step": pre_ms + choose_ms + local_ms,
               "evals_per_s": n_eval / (pre_ms * 1e-3),
               "restarts_per_s": (cfg.p_maxpoints / (local_ms * 1e-3)) if local_ms > 0 else None,
               "mean_evals_per_restart": float(np.mean(s.evals)) if s.evals is not None else None,
               "evals_in_restarts": float(np.sum(s.evals)) if s.evals is not None else None,
               "local_launches": getattr(s, "local_launches", None), "gpu_launches": int(launches),
               # the restart kernels alone (events around the launches, last step); stage_ms.local also holds the fetch of every row
               # (searchStart, searchResults, evaluation counts) into host arrays, which is what the caller gets back
               "local_kernels_ms": None if self.no_local else float(np.mean([d["local_kernels_ms"] for d in dev])),
               "rescue_calls": int(np.sum(s.nrescue)) if (self.name == "C5" and getattr(s, "nrescue", None) is not None) else None,  # RESCUE_H entries (DFNLS)
               "local_mode": "async" if self.local_async else "rounds",
               "instances": getattr(s, "instances_used", None) if self.local_async else None,
               "instance_form": (("one warp (nm_w4_kernel), tick = objective evaluation" if self.alg == "a" and s.asyncClock(self.alg, ASYNC_INSTANCES.get(self.name, 256)) == 1
                                  else "four-warp block (nm_quad_async_kernel), tick = loop trip" if self.alg == "a" else "block (dfnls_kernel), tick = objective evaluation")
                                 if (self.local_async and not self.no_local) else None),
               "pretest_ms_steps": [round(float(d["pretest_ms"]), 5) for d in dev],
               "best_fval": None if self.no_local else s.getBestLine()[0]}
        if self.local_async and not self.no_local:  # the same restarts as rounds of round_size, beside (untimed steps)
            self.local_async = False
            rd = [self.step(flush, False) for _ in range(3)][1:]  # (the first pass loads the rounds kernels: CUDA loads modules lazily)
            self.local_async = True
            out["rounds"] = {"local_ms": float(np.mean([d["local_ms"] for d in rd])), "local_kernels_ms": float(np.mean([d["local_kernels_ms"] for d in rd])), "local_launches": getattr(s, "local_launches", None),
                             "restarts_per_s": cfg.p_maxpoints / (np.mean([d["local_ms"] for d in rd]) * 1e-3), "best_fval": s.getBestLine()[0]}
        if self.local_async and not self.no_local and self.alg == "a" and cfg.nx <= 31 and "one warp" in (out["instance_form"] or ""):
            # ... and as four-warp blocks (the latency form: a third of the instances, each ~2x as fast), beside (untimed steps)
            fw = []
            for _ in range(3):
                flush.zero_(); self.barrier()
                s.solveAtSobolPoints(wait=False); s.chooseSobol()
                ea = self.mark(); s.LocalMinimizationsAsync(self.alg, instances=6 * 148); eb = self.mark()
                self.ext.synchronize()
                fw.append(ea.elapsed_time(eb))
                fwk = s.stage_ms("local")
            fw = fw[1:]
            out["four_warp_blocks"] = {"local_ms": float(np.mean(fw)), "local_kernels_ms": float(fwk), "instances": s.instances_used, "restarts_per_s": cfg.p_maxpoints / (np.mean(fw) * 1e-3),
                                       "best_fval": s.getBestLine()[0]}
        if self.local_async and not self.no_local and self.alg == "a":  # ... and as free-running instances (not reproducible: beside, untimed steps)
            fr = []
            for _ in range(3):
                flush.zero_(); self.barrier()
                s.solveAtSobolPoints(wait=False); s.chooseSobol()
                ea = self.mark(); s.LocalMinimizationsAsync(self.alg, instances=ASYNC_INSTANCES[self.name], free_running=True); eb = self.mark()
                self.ext.synchronize()
                fr.append(ea.elapsed_time(eb))
                frk = s.stage_ms("local")
            fr = fr[1:]  # (first pass: lazy module load)
            out["free_running"] = {"local_ms": float(np.mean(fr)), "local_kernels_ms": float(frk), "restarts_per_s": cfg.p_maxpoints / (np.mean(fr) * 1e-3), "best_fval": s.getBestLine()[0],
                                   "note": "no virtual clock: the reference's multi-process semantics literally; verified restart by restart (tests), not reproducible as a whole"}
        if e2e:
            out["e2e_sobol_s"] = self.mx(np.mean([d["wall_pretest"] for d in e2e]))
            out["e2e_step_s"] = self.mx(np.mean([d["wall"] for d in e2e]))
            out["e2e_copy_tail_s"] = self.mx(np.mean([d["copy_tail"] for d in e2e]))
            out["e2e_wall_local_s"] = self.mx(np.mean([d["wall_local"] for d in e2e]))
            out["d2h_total"] = int(self.mx(e2e[-1]["d2h"]) * self.world)
        return out


def rooflines(name, cfg, m, world, peak_tf, hbm_gbs, hbm_src):
    """roofline of the Sobol-stage kernel (the stage BASELINE.json sets the >= 50 % target on) and of the restart kernel"""
    per_gpu = m["n_evaluated"] / world
    eval_ms, step_ms = m["stage_ms"]["pretest_eval_kernel"], m["ms_per_step"]
    traffic, tsrc = TRAFFIC.get(name, (None, None)) if world == 1 else (None, None)
    if name == "C5":
        sm = cfg.extra["smm"]
        bpe = 8 * (3 * sm["np"] * sm["T"] + sm["np"]) + 8 * cfg.nm
        ach = per_gpu * bpe / (eval_ms * 1e-3) / 1e9
        fpe = 60.0 * sm["np"] * sm["T"]  # ~60 flops per person-age
        ach_tf = per_gpu * fpe / (eval_ms * 1e-3) / 1e12
        roof = {"bound": "hbm", "kernel": "smm_wave_kernel", "achieved": ach, "peak": hbm_gbs, "unit": "GB/s", "frac": ach / hbm_gbs,
                "traffic": traffic, "traffic_source": tsrc, "bytes_per_eval": bpe, "evals_per_launch": 8 * 148, "evals_per_step": per_gpu,
                "kernel_ms": eval_ms, "peak_source": hbm_src,
                "fp64": {"achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "flops_per_eval": fpe,
                         "note": "the BINDING limit: the 148 concurrent blocks walk the 127 MB panel together and it is served from the 126 MB L2 "
                                 "(ncu: hit rate 96 %, DRAM traffic 0.7 % of the algorithmic bytes), so the kernel is limited by the FP64 pipe "
                                 "(ncu: 62 % busy) together with shared-memory wavefronts (60 % of peak), not by HBM"},
                "note": "algorithmic bytes = the panel streamed once per evaluation (SURVEY 8d)",
                "share_of_step": eval_ms / step_ms if step_ms > 0 else None}
    else:
        ach = per_gpu * F_ALG[name] / (eval_ms * 1e-3) / 1e12
        roof = {"bound": "fp64", "kernel": "sobol_eval_kernel", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic, "traffic_source": tsrc, "flops_per_eval": F_ALG[name], "evals_per_launch": per_gpu, "kernel_ms": eval_ms,
                "peak_source": "measured: DFMA-chain microbenchmark (tiktak_cuda_fp64_peak) on this GPU; MEASURED_PEAKS.json has no FP64 entry",
                "share_of_step": eval_ms / step_ms if step_ms > 0 else None,
                "note": "CUDA-event duration: an EMPTY kernel between the same two events takes 6.3-7.5 us on this GPU "
                        "(tools/micro/launch_floor.cu), which is inside kernel_ms"}
    local = None
    if m.get("evals_in_restarts") and m["stage_ms"]["local"] > 0 and name != "C5":
        ach = m["evals_in_restarts"] * F_ALG[name] / (m["stage_ms"]["local"] * 1e-3) / 1e12
        asy = m.get("local_mode") == "async"
        w1 = asy and "one warp" in (m.get("instance_form") or "")
        local = {"bound": "latency", "kernel": (("nm_w4_kernel (one warp per restart, %s asynchronous instances, one launch)" if w1 else
                                                 "nm_quad_async_kernel (four warps per restart, %s asynchronous instances, one launch)") % m.get("instances")) if asy else
                 ("nm_quad_kernel (four warps per restart)" if cfg.nx <= 31 else "nm_w4_kernel (one warp per restart, device-side restart queue)"),
                 "achieved": ach, "peak": peak_tf * world, "unit": "TFLOP/s", "frac": ach / (peak_tf * world),
                 "objective_evals_per_step": m["evals_in_restarts"], "share_of_step": m["stage_ms"]["local"] / step_ms,
                 "note": ("every instance takes the next restart the moment it finishes one (virtual-time order, DESIGN.md section 3); a restart is a chain of "
                          "dependent simplex iterations, ~0.6-1.2 us per counted evaluation") if asy else
                         ("rounds are sequential (a round's starts depend on the best row of the rounds before it) and a round lasts as long "
                          "as its longest restart: a chain of dependent simplex iterations, ~0.6-1.2 us per counted evaluation (DESIGN.md section 3)")}
    return roof, local


def parity_block(rank, world, local_rank):
    """a small job sharded over the ranks of THIS run against the single-instance CPU oracle (rank 0 compares); both the
    no-cut and the legitimate-cut paths of the Sobol stage, selection and every restart -- outside any timed region"""
    from tiktak_b200 import OBJ_GRIEWANK, OBJ_RASTRIGIN, SOBOL_BLOCK, TikTakSearch, make_config

    cases = {
        "griewank10_nocut": dict(nx=10, nm=1, qr_ndraw=5 * SOBOL_BLOCK + 777, maxpoints=63, lo=-400.0, hi=600.0, init=[100.0] * 10,
                                 objective_id=OBJ_GRIEWANK, round_size=16),
        "rastrigin20_cut": dict(nx=20, nm=1, qr_ndraw=9 * SOBOL_BLOCK + 5, maxpoints=40, lo=-4.12, hi=5.12, init=[1.0] * 20,
                                objective_id=OBJ_RASTRIGIN, round_size=7, fvalmax=140000.0, legitimate=int(0.34 * (9 * SOBOL_BLOCK + 5))),
    }
    ok, detail = True, {}
    for cname, kw in cases.items():
        cfg = make_config(**kw)
        with TikTakSearch(cfg, device=local_rank, rank=rank, world=world) as s:
            _, ne, _ = s.solveAtSobolPoints()
            xs = s.chooseSobol()
            nl = s.n_legit
            s.LocalMinimizations("a")
            res, st, ev, best = s.searchResults, s.searchStart, s.evals, s.getBestLine()
        if rank == 0:
            from oracle.oracle import MATH_TK, Oracle

            o = Oracle(cfg, MATH_TK)
            _, one, onl = o.solveAtSobolPoints()
            oxs = o.chooseSobol(mode=1)
            o.LocalMinimizations(1)
            checks = {"cut": (ne, nl) == (one, onl), "x_starts": bool(np.array_equal(xs, oxs)), "starts": bool(np.array_equal(st, o.searchStart)),
                      "results": bool(np.array_equal(res, o.searchResults)), "evals": bool(np.array_equal(ev, o.evals)),
                      "best": best[0] == o.getBestLine()[0]}
            detail[cname] = checks
            ok = ok and all(checks.values())
        if world == 1:  # the asynchronous-instances form of the same restarts against the oracle's event simulation
            with TikTakSearch(cfg, device=local_rank) as s:
                if s.asyncSupported("a"):
                    s.solveAtSobolPoints()
                    s.chooseSobol()
                    s.LocalMinimizationsAsync("a")
                    res, st, ev, best, used = s.searchResults, s.searchStart, s.evals, s.getBestLine(), s.instances_used
                    o.LocalMinimizationsAsync(used, clock=s.asyncClock("a"))
                    checks = {"starts": bool(np.array_equal(st, o.searchStart)), "results": bool(np.array_equal(res, o.searchResults)),
                              "evals": bool(np.array_equal(ev, o.evals)), "best": best[0] == o.getBestLine()[0], "instances": int(used)}
                    detail[cname + "_async"] = checks
                    ok = ok and all(bool(v) for v in checks.values())
    return {"ok": bool(ok), "world": world, "checked": "Sobol-stage cut and valid count, x_starts, searchStart, searchResults (bit for bit), evaluation counts, best row; "
            "sharded over this run's ranks vs the single-instance CPU oracle", "cases": detail}


def run_graft(args, rank, world, local_rank):
    import ctypes as C

    import torch

    from tiktak_b200 import lib

    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = lib.load()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    name = args.workload
    cfg = workload_config(name, world, args.round_size, args.nm_shrink_mode)
    steps, warmup = max(1, args.steps), max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    sampler.start()
    job = Job(name, cfg, rank, world, local_rank, dist, args.no_local, args.local_mode)
    m = job.measure(flush, steps, warmup, sampler, e2e_steps=steps)
    job.close()

    fl, ms = C.c_double(), C.c_double()
    lib.check(L.tiktak_cuda_fp64_peak(local_rank, C.byref(fl), C.byref(ms)), "fp64_peak")
    peak_tf = fl.value / 1e12
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_gbs = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"

    per_workload, strong, parity = None, None, None
    if not args.no_extra:
        if world == 1:
            per_workload = {}
            for wn in ("C3", "C4", "C5"):
                if wn == name:
                    continue
                wcfg = workload_config(wn, 1)
                wj = Job(wn, wcfg, rank, 1, local_rank, None, local_mode=args.local_mode)
                wm = wj.measure(flush, 1 if wn == "C5" else 2, 1, sampler)
                wj.close()
                roof, rl = rooflines(wn, wcfg, wm, 1, peak_tf, hbm_gbs, hbm_src)
                wm.update(roofline=roof, roofline_local=rl,
                          cpu_baseline=None if args.no_cpu else cpu_arm(wn, 4.0, faithful=False))
                per_workload[wn] = wm
                if wn == "C3":
                    strong = dict(wm, n_gpus=1)
            # a dimension WITHOUT a compiled-in instantiation of the Sobol kernel: the library compiles the fixed-nx kernel for it
            # at run time (NVRTC, csrc/rtc.cu); Sobol stage + selection only
            from tiktak_b200 import OBJ_GRIEWANK, make_config
            xcfg = make_config(7, 1, 1_000_000, 1000, -400.0, 600.0, init=[100.0] * 7, objective_id=OBJ_GRIEWANK, round_size=128)
            xj = Job("X7", xcfg, rank, 1, local_rank, None, no_local=True)
            xm = xj.measure(flush, 3, 3, sampler)
            xj.close()
            xroof, _ = rooflines("X7", xcfg, xm, 1, peak_tf, hbm_gbs, hbm_src)
            L.tiktak_cuda_debug_rtc_kernels.restype = C.c_int
            xroof["kernel"] = "sobol_eval_kernel<ObjGriewank, 7, ...> compiled at run time" if L.tiktak_cuda_debug_rtc_kernels() > 0 else "sobol_eval_generic_kernel (runtime nx)"
            xm.update(roofline=xroof, rtc_kernels_loaded=int(L.tiktak_cuda_debug_rtc_kernels()))
            per_workload["X7"] = xm
        else:
            scfg = workload_config("C3", world, weak=False)
            sj = Job("C3", scfg, rank, world, local_rank, dist, local_mode=args.local_mode)
            sm_ = sj.measure(flush, 2, 1, sampler)
            sj.close()
            strong = dict(sm_, n_gpus=world)
        parity = parity_block(rank, world, local_rank)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank == 0:
        roof, rl = rooflines(name, cfg, m, world, peak_tf, hbm_gbs, hbm_src)
        cpu = cpu_arm(name, args.cpu_seconds) if (world == 1 and not args.no_cpu) else None
        tables_bytes = 32 * 4 * cfg.nx + 5 * 8 * cfg.nx + 64
        line = {
            "metric": "sobol_objective_evals_per_s", "value": m["evals_per_s"], "unit": "evals/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(name, cfg),
                       "parallelism": f"sobol index blocks + restarts sharded over {world} GPU(s); exchanges on the library's own NCCL communicator",
                       "l2": "256 MB buffer rewritten between steps (L2 flush)",
                       "timing": "CUDA events on the library stream per stage, max over ranks (pretest: the library's events around the "
                                 "stage's own device work; choose/local: events bracketing the host calls)"},
            "restarts_per_s": m["restarts_per_s"], "mean_evals_per_restart": m["mean_evals_per_restart"],
            "stage_ms": m["stage_ms"], "pretest_ms_steps": m["pretest_ms_steps"], "local_launches": m["local_launches"],
            "local_mode": m["local_mode"], "instances": m["instances"], "rounds": m.get("rounds"), "free_running": m.get("free_running"),
            "roofline": roof, "roofline_local": rl, "cpu_baseline": cpu,
            "e2e": {"value": m["n_evaluated"] / m["e2e_sobol_s"], "unit": "evals/s",
                    "h2d_bytes_per_step": tables_bytes, "d2h_bytes_per_step": m["d2h_total"],
                    "step_s": m["e2e_step_s"], "copy_tail_s": m["e2e_copy_tail_s"],
                    "restarts_per_s": (cfg.p_maxpoints / m["e2e_wall_local_s"]) if m["e2e_wall_local_s"] > 0 else None,
                    "what": "wall clock of solveAtSobolPoints through the public API with the stage's counters (n_evaluated, n_legit = lastSobol.dat / "
                            "legitSobol.dat) read back to the host every step; every stored value (the fval column of sobolFnVal.dat, d2h_bytes_per_step) "
                            "is streamed to pinned host memory on a second stream beside selection and the restarts and has landed before step_s "
                            "ends (copy_tail_s = what was still outstanding when the restarts were done)"},
            "gpu_launches": m["gpu_launches"], "clocks": sampler.summary(), "best_fval": m["best_fval"],
            "per_workload": per_workload, "strong_c3": strong, "parity": parity,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun (WORLD_SIZE={world})"}))
        sys.exit(2)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_graft(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
