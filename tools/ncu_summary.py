#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + per-source-line shares (needs -lineinfo and --import-source)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (units[i], vals[i]) for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tc.sum',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg']
print("| metric | value |\n|---|---|")
for k in keys:
    if k in m:
        print("| `%s` | %s %s |" % (k, m[k][1], m[k][0]))
for h in ('sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
          'smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed'):
    if h in m:
        print("| `%s` | %s %s |" % (h, m[h][1], m[h][0]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
lines = []
for r in csv.reader(src.splitlines()):
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    elif r and r[0].isdigit() and len(r) > 7:
        try:
            lines.append((cur, int(r[0]), r[1].strip(), int(r[4]) if r[4] not in ('-', '') else 0, int(r[7]) if r[7] not in ('-', '') else 0))
        except ValueError:
            pass
ts, ti = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
print("\nper-line shares (stall samples / instructions), lines above %.1f%%:\n" % (100 * thr))
for f, ln, s, sm, ins in sorted(lines, key=lambda x: (x[0], x[1])):
    if sm / ts > thr or ins / ti > thr:
        print("%-14s %4d %6.2f%% smp %6.2f%% inst  %s" % (f, ln, 100 * sm / ts, 100 * ins / ti, s[:110]))
