#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): key raw metrics + per-opcode executed-instruction split + top stall sites."""
import csv, collections, re, subprocess, sys
rep = sys.argv[1]; tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'local_load', 'smsp__inst_executed_op_local']
for h, u, v in zip(hdr, units, vals):
    if h in keep or h.startswith('smsp__average_warps_issue_stalled') and float(v or 0) > 0.15:
        print("%s,%s,%s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]; ix = h.index("Instructions Executed"); isrc = h.index("Source"); ismp = h.index("# Samples")
ops = collections.Counter(); smp = collections.Counter(); tot = 0; lines = []
for r in rows[2:]:
    if len(r) <= ix: continue
    s = r[isrc].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s)
    op = '.'.join((m.group(2) if m else s).split('.')[:2])
    n = int(r[ix]); ops[op] += n; tot += n; smp[op] += int(r[ismp]); lines.append((int(r[ismp]), n, s))
print("total warp instructions", tot, "per tile", tot / tiles)
for op, n in ops.most_common(28):
    print("  %-20s %12d %6.2f%% per-tile %8.0f samples %d" % (op, n, 100 * n / tot, n / tiles, smp[op]))
print("top sampled instructions:")
for sm, n, s in sorted(lines, reverse=True)[:25]:
    print("  samples %6d execs %10d  %s" % (sm, n, s[:110]))
