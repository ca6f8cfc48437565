#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + executed instructions by opcode + top stall lines.
usage: ncu_summary.py report.ncu-rep [kernel-index]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, rows = r[0], r[1], r[2:]
print("kernels:", [row[hdr.index('Kernel Name')][:40] for row in rows])
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print("%-80s %s %s" % (w, [row[i] for row in rows], units[i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; data = []; k = 0
for row in rows:
    if row and row[0] == "Address":
        hdr = row; k += 1; continue
    if hdr and k == 1 and len(row) > 5:
        data.append(row)
if hdr:
    ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    byop, samp, tot, stot = collections.Counter(), collections.Counter(), 0, 0
    for row in data:
        try:
            n = int(row[ia]); sm = int(row[ist])
        except ValueError:
            continue
        t = row[isrc].split()
        op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '?')).split('.')[0]
        byop[op] += n; samp[op] += sm; tot += n; stot += sm
    print("SASS lines %d, warp-instructions executed %d, samples %d" % (len(data), tot, stot))
    for op, n in byop.most_common(16):
        print("  %-10s %14d %5.1f%%   stall-samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(stot, 1)))

# ---- executed instructions / stall samples per CUDA source line (needs -lineinfo + --import-source on)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; hdr = None; lines = []
for row in rows:
    if row and row[0] in ("File Path", "File Name"): cur = row[1].split('/')[-1]; continue
    if row and row[0] == "Line No": hdr = row; continue
    if hdr and len(row) > 8 and row[0].isdigit():
        try:
            lines.append((int(row[hdr.index("Instructions Executed")]), int(row[hdr.index("# Samples")]), cur, int(row[0]), row[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1; stot = sum(l[1] for l in lines) or 1
print("top source lines by executed warp-instructions:")
for n, sm, f, ln, txt in sorted(lines, reverse=True)[:28]:
    print("  %5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * n / tot, 100.0 * sm / stot, f, ln, txt))

# ---- the same, grouped by region of the sweep kernel (line ranges of rvn_engine.cu / whole helper files)
def region(f, ln):
    if f == "rvn_engine.cu":
        for name, lo, hi in (("forcings", 29, 191), ("conv pipeline+div", 194, 264), ("conv loop", 265, 362), ("block_sum", 363, 377),
                             ("prologue", 378, 485), ("process interpreter", 486, 512), ("epilogue", 513, 571)):
            if lo <= ln <= hi:
                return name
        return "engine other"
    if f == "rvn_processes.cuh":
        return "process rate functions"
    return f
reg = collections.Counter(); regs = collections.Counter()
for n, sm, f, ln, txt in lines:
    reg[region(f, ln)] += n; regs[region(f, ln)] += sm
print("by region (line ranges valid for the source revision profiled):")
for r, n in reg.most_common():
    print("  %-28s %5.1f%% inst %5.1f%% stall" % (r, 100.0 * n / tot, 100.0 * regs[r] / stot))
