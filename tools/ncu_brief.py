#!/usr/bin/env python3
"""Brief of one ncu report: key metrics, stall mix, top stalled SASS instructions.  usage: tools/ncu_brief.py rep [topn] [kernel-regex]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 15
kf = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + kf, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct']
for r in rows[2:]:
    d = dict(zip(h, r))
    print("kernel:", d.get("Kernel Name", "")[:60])
    for k in keys: print('  %-70s %s' % (k, d.get(k)))
    st = {k: float(v.replace(',', '')) for k, v in d.items() if k.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in k and v not in ('', 'n/a')}
    tot = sum(st.values()) or 1
    print('  stalls: ' + ' | '.join('%s %.1f%%' % (k.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + kf, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ia = h.index("Instructions Executed"); isrc = h.index("Source"); isamp = h.index("# Samples")
data = [r for r in rows[2:] if len(r) == len(h) and r[ia].isdigit()]
tot = sum(int(r[isamp]) for r in data) or 1
for i in sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]:
    r = data[i]; print("%5.2f%% exec=%-10s idx=%-6d %s" % (100 * int(r[isamp]) / tot, r[ia], i, r[isrc].strip()[:90]))
