#!/usr/bin/env python
"""Text summary of every kernel in one or more .ncu-rep files (ncu --set full): duration, DRAM bytes and achieved GB/s against the
measured HBM peak, tensor / issue / XU / L1 / L2 utilisation, registers, top stall reasons.  No GPU needed."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM = peaks["hbm_gbs"]
M = {"t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
     "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "regs": "launch__registers_per_thread", "inst": "smsp__inst_executed.sum", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active"}
def unit_scale(u):
    return {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
print("measured HBM copy peak %.0f GB/s (MEASURED_PEAKS.json); one capture per kernel under ncu --set full --clock-control none (cold caches, serialised)" % HBM)
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    col = {name: i for i, name in enumerate(h)}
    print("== %s" % os.path.basename(rep))
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        g = lambda k: float(r[col[M[k]]] or 0) * unit_scale(u[col[M[k]]]) if M[k] in col else float("nan")
        t_us, rd, wr = g("t"), g("rd"), g("wr")
        gbs = (rd + wr) / (t_us * 1e-6) / 1e9 if t_us else 0.0
        stalls = sorted(((float(r[i] or 0), h[i][len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for i in range(len(h))
                         if h[i].startswith("smsp__average_warps_issue_stalled_") and h[i].endswith("_per_issue_active.ratio")), reverse=True)[:4]
        print("%-46s %8.1f us  dram rd %7.1f MB wr %7.1f MB = %6.0f GB/s (%.2f of peak) | tensor %4.1f%% issue %4.1f%% xu %4.1f%% l1 %4.1f%% l2 %4.1f%% | regs %3d warps %4.1f%% | winst %.1fM | stalls %s"
              % (name[:46], t_us, rd / 1e6, wr / 1e6, gbs, gbs / HBM, g("tensor"), g("issue"), g("xu"), g("l1"), g("l2"), int(g("regs")), g("warps"), g("inst") / 1e6,
                 ", ".join("%s %.1f" % (n, v) for v, n in stalls)))
