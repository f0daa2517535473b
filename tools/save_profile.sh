#!/bin/bash
# usage: tools/save_profile.sh <tag> <bench.json> <launches.csv> <report.ncu-rep>   -> profiles/<tag>_*
set -e
tag=$1
cp "$2" profiles/${tag}_bench.json
cp "$3" profiles/${tag}_launches.csv
ncu -i "$4" --page details --csv > profiles/${tag}_ncu_details.csv
ncu -i "$4" --page raw --csv > profiles/${tag}_ncu_raw.csv
python tools/ncu_lines.py "$4" 40 > profiles/${tag}_stalls_by_line.txt 2>&1
ncu -i "$4" --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum.per_second','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d=dict(zip(h,r)); du=dict(zip(h,u))
    print('kernel:', d['Kernel Name'][:70])
    for k in keys: print('  %-72s %s %s'%(k,d.get(k),du.get(k)))
    st={k:float(v.replace(',','')) for k,v in d.items() if k.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in k and v not in ('','n/a')}
    tot=sum(st.values()) or 1
    print('  stalls: '+' | '.join('%s %.1f%%'%(k.replace('smsp__pcsamp_warps_issue_stalled_',''),100*v/tot) for k,v in sorted(st.items(),key=lambda x:-x[1])[:10]))
" > profiles/${tag}_summary.txt
cat profiles/${tag}_summary.txt
