"""Summarise an .ncu-rep: key raw metrics per kernel + dynamic instruction mix + top stall lines.  usage: ncu_summary.py rep [--mix] [--top N]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'inst_executed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-90s %s %s" % (w, r[i][:100], units[i]))
    print('---')
if '--mix' in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    iA, iE, iS = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    c, s = Counter(), Counter(); tot = 0; data = []
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name': break
        if len(r) < len(hdr) or not r[0].startswith('0x'): continue
        a = r[iA].strip(); e = int(r[iE]); sm = int(r[iS])
        op = a.split()[0]
        if op.startswith('@'): op = a.split()[1]
        op = op.split('.')[0].rstrip(';')
        c[op] += e; s[op] += sm; tot += e; data.append((a, e, sm))
    print("total warp-instructions %d, stall samples %d" % (tot, sum(s.values())))
    for op, e in c.most_common(28): print("  %-8s %10.2f M  samples %d" % (op, e / 1e6, s[op]))
    n = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 0
    if n:
        print("top stall instructions:")
        for i, (a, e, sm) in sorted(enumerate(data), key=lambda t: -t[1][2])[:n]:
            print("  #%d samples %d exec %d : %s" % (i, sm, e, a))
