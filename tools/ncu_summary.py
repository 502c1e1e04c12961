#!/usr/bin/env python3
"""Turn ncu output brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_c2.csv profiles/r01_launches_c2.csv "<note>"
        per-kernel launch count / total / mean duration and share of the captured region, from the
        `ncu --metrics gpu__time_duration.sum --csv --log-file` launch list
    python tools/ncu_summary.py full gpurun_out/prof_c2.ncu-rep profiles/r01_full_c2.csv "<note>"
        the handful of `--set full` counters DESIGN.md quotes, one column per captured launch
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
]


def launches(src, dst, note):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    # not part of a step: the FP64-peak microbenchmark (roofline denominator) and torch's L2-flush fill between steps
    outside = {k: v for k, v in agg.items() if "dfma_chain" in k or "vectorized_elementwise" in k}
    agg = OrderedDict((k, v) for k, v in agg.items() if k not in outside)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {note}\n# source: {src} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches)\n")
        f.write("kernel,launches,total_us,mean_us,share_of_step_kernels\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{ns / 1e3:.2f},{ns / 1e3 / n:.2f},{ns / tot:.4f}\n")
        f.write(f"TOTAL,{sum(a[0] for a in agg.values())},{tot / 1e3:.2f},,1.0\n")
        for k, (n, ns) in outside.items():
            f.write(f"# outside the timed step: \"{k}\",{n},{ns / 1e3:.2f}\n")


def full(src, dst, note):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# {note}\n# source: {src} (ncu --set full --clock-control none --import-source on)\n")
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(data))) + "\n")
        f.write("Kernel Name,," + ",".join('"' + r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "") + '"' for r in data) + "\n")
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"{m},{units[i]}," + ",".join(r[i].replace(",", "") for r in data) + "\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    {"launches": launches, "full": full}[mode](src, dst, note)
