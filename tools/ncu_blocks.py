"""Summarises an `ncu --page source --print-source sass --csv` dump: blocks of consecutive SASS instructions with the
same execution count, their share of all executed warp-instructions and their average active lanes.
Usage: python tools/ncu_blocks.py raw.csv sass.csv"""
import csv
import sys

raw, sass = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr = rows[0]
keys = ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum')
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get('Kernel Name', '')[:70])
    for k in keys:
        print('  ', k, d.get(k))
rows = list(csv.reader(open(sass)))
hdr = rows[1]
data = rows[2:]
ia, ith, isrc, ismp = (hdr.index(x) for x in ('Instructions Executed', 'Avg. Threads Executed', 'Source', '# Samples'))
tot = sum(int(r[ia]) for r in data)
prev, blk, out = None, [], []
for k, r in enumerate(data):
    c = int(r[ia])
    if prev is None or abs(c - prev) > 0.02 * max(c, prev, 1):
        if blk:
            out.append(blk)
        blk = []
    blk.append((k, c, r[isrc].strip(), r[ith], r[ismp]))
    prev = c
out.append(blk)
for blk in out:
    n, c = len(blk), blk[0][1]
    smp = sum(int(b[4]) for b in blk)
    if n * c / max(tot, 1) > 0.01:
        print(f"rows {blk[0][0]:4d}-{blk[-1][0]:4d} n={n:3d} exec/inst={c:9d} ({100*n*c/tot:5.1f}%) lanes={blk[0][3]:>5s} samples={smp}  {blk[0][2][:40]}")
print('total warp-instructions', tot)
