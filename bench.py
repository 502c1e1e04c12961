#!/usr/bin/env python3
"""bench.py -- TikTak hot path on B200: Sobol objective evals/s and local restarts/s.

    python bench.py --gpus N --steps K --warmup W [--workload C2|C3|C4] [--impl reference]

One "step" = one pass of the hot path over one synthetic job: Sobol pre-test of qr_ndraw points
(solveAtSobolPoints) -> start selection (chooseSobol) -> all Nelder-Mead restarts
(LocalMinimizations: rounds of round_size restarts, later rounds started early on idle SMs and kept
while the best row is unchanged -- same rows as strictly sequential rounds), device-resident.  N=1 runs BASELINE.json configs[1] (Griewank 10-D, 1e6 Sobol
points, 1000+1 restarts).  N>1 (torchrun, one rank per GPU, NCCL) shards the Sobol index range and
each round's restarts over the ranks with one all-gather per round; per-GPU work is held fixed
(qr_ndraw, maxpoints and round_size scale with N): "scaling": "weak".

JSON keys beyond the base contract:
  restarts_per_s   local restarts/s of the same step (second half of BASELINE.json's metric)
  roofline_local   the restart kernel (dominant in the step's kernel time, latency-bound): algorithmic FP64
                   flops of its objective evaluations / stage time, against the same FP64 peak
  roofline         the Sobol evaluation kernel (the stage BASELINE.json sets the >=50 % target on):
                   algorithmic FP64 flops per evaluation x evaluations per launch /
                   CUDA-event duration on the library's stream, vs the FP64 DFMA peak measured on
                   this box by tiktak_cuda_fp64_peak (MEASURED_PEAKS.json carries no FP64 figure)
  cpu_baseline     the oracle port timed on this box's host cores on a bounded sample
  e2e              same metric through the public host API with HOST buffers (fvals copied back)
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
F_ALG = {"C1": 237, "C2": 545, "C3": 1105, "C4": 905}
# restarts per round (all blended against the best row of the round's start).  C4: every restart runs into the
# 5000-evaluation cap, so a round costs whole waves of blocks: 1184 = 148 SMs x 4 resident blocks x 2 waves
ROUND_SIZE = {"C1": 4, "C2": 128, "C3": 1024, "C4": 1184, "C5": 128}
# C5 (synthetic SMM): HBM-bound, the panel is streamed once per evaluation: 8*(3*np*T + np) + 8*nm bytes
C5_SIZES = dict(np_persons=131072, T=40, qr_ndraw=20000, legitimate=10000, maxpoints=127)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--round-size", type=int, default=0)
    ap.add_argument("--nm-shrink-mode", type=int, default=-1, choices=[-1, 0, 1],
                    help="initial-simplex rule: 0 = the reference code, 1 = README.md:36 (bounds from the best minima so far); "
                         "default 0 everywhere: the rule is not in the reference's code (SURVEY finding 3) and on C4 it ends at a "
                         "worse minimum (3.42 against 1.002), so the parity configuration keeps the code's behaviour")
    ap.add_argument("--no-local", action="store_true", help="time only the Sobol stage + selection")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    return ap.parse_args()


def workload_config(name, world, round_size=0, nm_shrink_mode=-1):
    from tiktak_b200 import baseline_config, smm_config

    cfg = smm_config(**C5_SIZES) if name == "C5" else baseline_config(name)
    cfg.nm_shrink_mode = 0 if nm_shrink_mode < 0 else nm_shrink_mode
    if world > 1:  # weak scaling: per-GPU work fixed
        cfg.qr_ndraw *= world
        cfg.maxpoints *= world
        cfg.legitimate = cfg.legitimate * world if name == "C5" else cfg.qr_ndraw + 10000
        cfg.lottery_points = cfg.maxpoints
    cfg.round_size = (round_size or ROUND_SIZE[name]) * world
    return cfg


# ------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons through NVML while the timed region runs"""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
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
    name, world, npts, nrestarts, round_size = args
    sys.path.insert(0, ROOT)
    from oracle.oracle import MATH_LIBM, Oracle

    cfg = workload_config(name, world, round_size)
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

    name, world, npts = args
    sys.path.insert(0, ROOT)
    from oracle.oracle import MATH_LIBM, Oracle

    cfg = workload_config(name, world, 0)
    cfg.legitimate = cfg.qr_ndraw + 10000
    o = Oracle(cfg, MATH_LIBM)
    o.L.orc_pretest_faithful.restype = C.c_long
    o.L.orc_pretest_faithful.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        n = o.L.orc_pretest_faithful(o.h, d.encode(), npts)
        t1 = time.perf_counter()
    return t1 - t0, int(n)


def cpu_arm(name, world, budget_s, round_size):
    """The reference cannot be compiled here (no Fortran compiler, ifort-only source); its CPU
    implementation of the path is timed as the line-faithful oracle port: P independent instances
    (the reference's own parallel model, README.md:30-33) on P host cores, compute-only (no .dat I/O)."""
    import multiprocessing as mp

    from oracle import oracle as orc

    orc.load()
    cores = os.cpu_count() or 1
    # calibrate on a tiny sample, then size the bounded sample to the budget
    n_cal = 3 if name == "C5" else 2000
    t_s, t_l, n0, r0, _ = _cpu_worker((name, world, n_cal, 2, round_size))
    per_eval, per_restart = t_s / n0, (t_l / r0 if r0 else 1.0)
    npts = int(max(n_cal, min(2_000_000, 0.5 * budget_s / per_eval)))
    nres = int(max(2, min(2000, 0.5 * budget_s / per_restart)))
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        outs = pool.map(_cpu_worker, [(name, world, npts, nres, round_size)] * cores)
        wall = time.perf_counter() - t0
    ts = max(o[0] for o in outs)
    tl = max(o[1] for o in outs)
    evals_s = cores * npts / ts
    restarts_s = (cores * (nres + 1) / tl) if tl > 0 else evals_s / 320.0
    faithful = None
    if name != "C5":  # what the reference's users pay: ~6 file open/close cycles around every evaluation
        with mp.get_context("fork").Pool(cores) as pool:
            nf = 20000
            outs_f = pool.map(_cpu_worker_faithful, [(name, world, nf)] * cores)
        faithful = {"value": cores * nf / max(o[0] for o in outs_f), "unit": "evals/s",
                    "sample": f"{cores} instances x {nf} Sobol points with the lastSobol/legitSobol/logFile/sobolFnVal "
                              f"file cycles of TiktakGlobalSearch.f90:426-463 in a private temp directory"}
    return {"value": evals_s, "unit": "evals/s", "cores": cores, "kind": "port", "with_file_protocol": faithful,
            "sample": f"{cores} independent oracle instances x ({npts} Sobol points of {name} + {nres + 1} "
                      f"Nelder-Mead restarts), libm cos, compute-only; wall {wall:.1f}s",
            "sample_ms": 1e3 * (ts + tl), "restarts_per_s": restarts_s, "mean_evals_per_restart": float(np.mean([o[4] for o in outs])),
            "single_core_evals_per_s": npts / float(np.mean([o[0] for o in outs]))}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cb = cpu_arm(args.workload, world, max(args.cpu_seconds, 5.0) * max(1, min(args.steps, 3)) / 3.0, args.round_size)
    cfg = workload_config(args.workload, world, args.round_size)
    line = {"impl": "reference", "metric": "sobol_objective_evals_per_s", "value": cb["value"], "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["sample_ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: nx={cfg.nx}, qr_ndraw={cfg.qr_ndraw}, maxpoints={cfg.maxpoints}",
                       "note": "reference = CPU oracle port (Fortran reference not buildable here), bounded sample"},
            "restarts_per_s": cb["restarts_per_s"], "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- GPU arm
def run_graft(args, rank, world, local_rank):
    import ctypes as C

    import torch

    from tiktak_b200 import TikTakSearch, lib

    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    name = args.workload
    cfg = workload_config(name, world, args.round_size, args.nm_shrink_mode)
    L = lib.load()
    s = TikTakSearch(cfg, device=local_rank, rank=rank, world=world)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    from tiktak_b200 import SOBOL_BLOCK
    n_host = cfg.qr_ndraw + 1 if world == 1 else ((cfg.qr_ndraw // SOBOL_BLOCK + world) // world) * SOBOL_BLOCK
    pinned = torch.empty(n_host, dtype=torch.float64, pin_memory=True).numpy()  # host side of the e2e copy

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ext = s.torch_stream()  # the library's stream: CUDA events recorded on it time the stages on the device
    alg = "b" if name == "C5" else "a"

    def mark():
        e = torch.cuda.Event(enable_timing=True)
        e.record(ext)
        return e

    def one_step(e2e):
        """returns dict of device stage times (ms, CUDA events on the library stream) and wall times (s)"""
        out = {}
        flush.zero_()  # evict L2 between steps (untimed)
        barrier()
        t0 = time.perf_counter()
        e0 = mark()
        _, ne, nl = s.solveAtSobolPoints(wait=False)
        e1 = mark()
        if e2e:
            ext.synchronize()
        t1 = time.perf_counter()
        if e2e:
            # D2H of every evaluated fval (the fval column of sobolFnVal.dat) into pinned host memory; a shard of a
            # multi-GPU run copies the values it owns
            fv = s.sobolFnVal(out=pinned) if world == 1 else s.sobolFnValShard(out=pinned)
            out["d2h"] = fv.nbytes
        t2 = time.perf_counter()
        e2 = mark()
        s.chooseSobol()
        e3 = mark()
        t3 = time.perf_counter()
        if not args.no_local:
            s.LocalMinimizations(alg)
        e4 = mark()
        ext.synchronize()
        t4 = time.perf_counter()
        # the pre-test stage is timed by the library's own events (first device operation of the stage -> last; the
        # NCCL count reductions of a sharded run with a possible cut are queued on the same stream in between)
        out.update(pretest_ms=s.stage_ms("pretest"), pretest_bracket_ms=e0.elapsed_time(e1), choose_ms=e2.elapsed_time(e3),
                   local_ms=0.0 if args.no_local else e3.elapsed_time(e4), pretest_eval_ms=s.stage_ms("pretest_eval"),
                   wall_pretest=t1 - t0, wall_values=t2 - t1, wall_choose=t3 - t2, wall_local=t4 - t3, wall=t4 - t0,
                   ne=ne, nl=s.n_legit, d2h=out.get("d2h", 0))
        return out

    for _ in range(max(args.warmup, 3)):
        one_step(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = s.launch_count()
    dev_steps = [one_step(False) for _ in range(args.steps)]
    launches = (s.launch_count() - l0) // max(args.steps, 1)
    e2e_steps = [one_step(True) for _ in range(args.steps)]
    sampler.stop_flag = True
    sampler.join(timeout=2)

    def mx(v):  # max over ranks
        if dist is None:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    npts, nrest = cfg.qr_ndraw, cfg.p_maxpoints
    # CUDA events on the library's stream (the NCCL exchanges of a sharded run are queued on the same stream),
    # mean over the timed steps, max over ranks
    pre_ms = mx(np.mean([d["pretest_ms"] for d in dev_steps]))
    choose_ms = mx(np.mean([d["choose_ms"] for d in dev_steps]))
    local_ms = mx(np.mean([d["local_ms"] for d in dev_steps]))
    timing = ("CUDA events on the library stream per stage, max over ranks (pretest: the library's events around the "
              "stage's own device work; choose/local: events bracketing the host calls)")
    eval_ms = mx(np.mean([d["pretest_eval_ms"] for d in dev_steps]))
    step_ms = pre_ms + choose_ms + local_ms
    n_eval_total = s.n_evaluated if s.n_evaluated else npts  # the legitimate cut may stop early (C5)
    value = n_eval_total / (pre_ms * 1e-3)
    restarts_s = (nrest / (local_ms * 1e-3)) if local_ms > 0 else None
    e2e_sobol_s = mx(np.mean([d["wall_pretest"] + d["wall_values"] for d in e2e_steps]))
    e2e_step_s = mx(np.mean([d["wall"] for d in e2e_steps]))
    wall_local = mx(np.mean([d["wall_local"] for d in e2e_steps]))
    d2h_total = int(mx(e2e_steps[-1]["d2h"]) * world)

    if rank == 0:
        fl, ms = C.c_double(), C.c_double()
        lib.check(L.tiktak_cuda_fp64_peak(local_rank, C.byref(fl), C.byref(ms)), "fp64_peak")
        peak_tf = fl.value / 1e12
        per_gpu_evals = (s.n_evaluated if name == "C5" else npts) / world
        if name == "C5":
            sm = cfg.extra["smm"]
            bpe = 8 * (3 * sm["np"] * sm["T"] + sm["np"]) + 8 * cfg.nm
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
            ach = per_gpu_evals * bpe / (eval_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "smm_wave_kernel", "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                        "frac": ach / peak_gbs,
                        # dram__bytes_read.sum + dram__bytes_write.sum of one launch = 1184 evaluations (148 SMs x 8),
                        # profiles/r01c_full_c5_smm.csv: the 127 MB panel is served from the 126 MB L2 (hit rate 96 %)
                        "traffic": 1020513512 if (world == 1 and sm["np"] == 131072 and sm["T"] == 40) else None,
                        "bytes_per_eval": bpe, "evals_per_launch": 8 * 148, "evals_per_step": per_gpu_evals,
                        "kernel_ms": eval_ms, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                        "note": "algorithmic bytes = the panel streamed once per evaluation (SURVEY 8d); the 148 concurrent "
                                "blocks share the panel through the 126 MB L2, so DRAM traffic is ~0.7 % of this figure and "
                                "the kernel is in fact co-limited by the FP64 pipe (62 % busy) and shared-memory wavefronts "
                                "(60 % of peak), ncu",
                        "fp64_peak_tflops": peak_tf, "share_of_step": eval_ms / step_ms if step_ms > 0 else None}
        else:
            achieved_tf = per_gpu_evals * F_ALG[name] / (eval_ms * 1e-3) / 1e12
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed `ncu --set full` captures
            # (profiles/r01d_full_*.csv); the 8 MB of results of a 1e6-point launch stay in the 126 MB L2
            traffic = {"C2": 25088, "C3": 742264320, "C4": 21494528}.get(name) if world == 1 else None
            roofline = {"bound": "fp64", "kernel": "sobol_eval_kernel", "achieved": achieved_tf, "peak": peak_tf,
                        "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                        "flops_per_eval": F_ALG[name], "evals_per_launch": per_gpu_evals, "kernel_ms": eval_ms,
                        "peak_source": "measured: DFMA-chain microbenchmark (tiktak_cuda_fp64_peak) on this GPU; "
                                       "MEASURED_PEAKS.json has no FP64 entry",
                        "share_of_step": eval_ms / step_ms if step_ms > 0 else None,
                        "note": "CUDA-event duration: an EMPTY kernel between the same two events takes 6.3-7.5 us on this "
                                "GPU (tools/micro/launch_floor.cu, gpurun_out/launch_floor.txt), which is inside kernel_ms; "
                                "ncu's gpu__time_duration of the C2 launch is 26.8 us (profiles/r01d_full_c2_sobol.csv)"}
        # the local stage's kernel dominates the STEP (profiles/: ~95 % of the kernel time at C2) but is a chain of
        # dependent simplex iterations, not a throughput kernel: reported for completeness against the same FP64 peak
        roofline_local = None
        if s.evals is not None and local_ms > 0 and name != "C5":
            evals_total = float(np.sum(s.evals))
            ach = evals_total * F_ALG[name] / (local_ms * 1e-3) / 1e12
            roofline_local = {"bound": "latency", "kernel": "nm_quad_kernel" if cfg.nx <= 31 else "nm_quad_wide_kernel",
                              "achieved": ach, "peak": peak_tf * world, "unit": "TFLOP/s", "frac": ach / (peak_tf * world),
                              "objective_evals_per_step": evals_total, "share_of_step": local_ms / step_ms,
                              "note": "one restart per 4-warp block; a round lasts as long as its longest restart; each "
                                      "simplex iteration is ~930 dependent cycles (DESIGN.md section 3)"}
        cpu = cpu_arm(name, world, args.cpu_seconds, args.round_size) if (world == 1 and not args.no_cpu) else None
        tables_bytes = 32 * 4 * cfg.nx + 5 * 8 * cfg.nx + 64
        line = {
            "metric": "sobol_objective_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {['', 'template', 'Griewank', 'Rastrigin', 'Rosenbrock', 'synthetic SMM'][int(name[1])]} "
                                   f"{cfg.nx}-D, qr_ndraw={npts}, maxpoints={cfg.maxpoints} (+1 = {nrest} Nelder-Mead restarts), "
                                   f"round_size={cfg.round_size}, nm_itmax={cfg.nm_itmax}, ftol={cfg.nm_ftol}, nm_shrink_mode={cfg.nm_shrink_mode}",
                       "parallelism": f"sobol index blocks + restarts sharded over {world} GPU(s)",
                       "l2": "256 MB buffer rewritten between steps (L2 flush)", "timing": timing},
            "restarts_per_s": restarts_s,
            "mean_evals_per_restart": float(np.mean(s.evals)) if s.evals is not None else None,
            "stage_ms": {"pretest": pre_ms, "pretest_eval_kernel": eval_ms, "choose": choose_ms, "local": local_ms},
            "pretest_ms_steps": [round(float(d["pretest_ms"]), 5) for d in dev_steps],  # this rank's timed steps, one by one
            "roofline": roofline,
            "roofline_local": roofline_local,
            "cpu_baseline": cpu,
            "e2e": {"value": n_eval_total / e2e_sobol_s, "unit": "evals/s",
                    "h2d_bytes_per_step": tables_bytes, "d2h_bytes_per_step": d2h_total,
                    "step_s": e2e_step_s, "restarts_per_s": (nrest / wall_local) if wall_local > 0 else None,
                    "what": "solveAtSobolPoints + sobolFnVal (all fvals to host) through the public API, wall clock"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "best_fval": s.getBestLine()[0] if not args.no_local else None,
        }
        print(json.dumps(line))
    s.close()
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
