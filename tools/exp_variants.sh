#!/bin/bash
# usage (on the GPU box): tools/exp_variants.sh <blocks> <variant> [variant ...]
# per variants/<variant>.so: quick bench (value, decode ms), then DRAM bytes / L2 hit rate of one decode launch (ncu metrics only)
lib=srslte-emane_b200/libsrslte_b200.so
cp $lib /tmp/default_lib.so
n=$1; shift
for v in "$@"; do
  echo "== $v"; cp variants/$v.so $lib
  bash tools/bench_quick.sh $n
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:tdec_win -s 3 -c 1 --csv --log-file gpurun_out/exp_$v.csv \
      python bench.py --blocks $n --steps 1 --warmup 3 --no-cpu --e2e-blocks 1024 > gpurun_out/exp_$v.log 2>&1
  python - gpurun_out/exp_$v.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print("   ", d.get("Metric Name"), d.get("Metric Value"), d.get("Metric Unit"))
PY
done
cp /tmp/default_lib.so $lib
