#!/usr/bin/env python
"""Summarise an ncu report: headline metrics + per-source-line share of stall samples and instructions.
    python tools/ncu_lines.py report.ncu-rep [units_per_warp_instruction_denominator]
"""
import csv, subprocess, sys, io
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 and float(sys.argv[2]) > 0 else None
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # which profiled launch of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, r = rows[0], rows[2 + kidx]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want:
        print("%-75s %s %s" % (h, r[i], rows[1][i]))
st = []
for i, h in enumerate(hdr):
    if 'stall' in h and 'ratio' in h and 'not_issued' not in h:
        try:
            st.append((float(r[i].replace(',', '')), h))
        except ValueError:
            pass
for v, h in sorted(st, reverse=True)[:7]:
    print("%-75s %.3f" % (h, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the page is a sequence of per-file tables; a launch starts at each table of our own .cu file
secs = [i for i, x in enumerate(rows) if x and x[0] == "File Path"]
mine = [i for i in secs if rows[i][1].endswith(".cu")]
lo = mine[min(kidx, len(mine) - 1)]
hi = mine[kidx + 1] if kidx + 1 < len(mine) else len(rows)
h = rows[lo + 2]; iS = h.index('# Samples'); iI = h.index('Instructions Executed')
body, other = [], [0, 0]
cur_mine = True
for i in range(lo, hi):
    x = rows[i]
    if x and x[0] == "File Path":
        cur_mine = x[1].endswith(".cu")
        continue
    if len(x) > iI and x[0].isdigit() and x[iS].isdigit():
        if cur_mine:
            body.append(x)
        else:
            other[0] += int(x[iS]); other[1] += int(x[iI])
print("\n(samples / instructions attributed to inlined CUDA header lines: %d / %d)" % tuple(other))
tot = sum(int(x[iS]) for x in body); totI = sum(int(x[iI]) for x in body)
agg = {}
for x in body:
    k = int(x[0]); a = agg.setdefault(k, [0, 0, x[1]]); a[0] += int(x[iS]); a[1] += int(x[iI])
print("\nline   samples  instr" + ("  instr/unit" if units else "") + "   source   (total warp instr %d)" % totI)
for k in sorted(agg):
    s_, i_, srcl = agg[k]
    if i_ / totI > 0.01 or s_ / tot > 0.012:
        extra = (" %6.2f" % (i_ / units)) if units else ""
        print("%5d %6.1f%% %6.1f%%%s | %s" % (k, 100 * s_ / tot, 100 * i_ / totI, extra, srcl.strip()[:100]))
