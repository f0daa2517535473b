#!/bin/bash
# usage: tools/bench_quick.sh <blocks> [more blocks...]   -- prints value, decode ms, fallbacks
for n in "$@"; do
  timeout 200 python bench.py --blocks $n --e2e-blocks 1024 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > /tmp/bq.json
  python - "$n" <<'PY'
import json,sys
d=json.loads(open('/tmp/bq.json').read().strip().splitlines()[-1])
print("blocks", sys.argv[1], "Gbit/s %.2f" % d["value"], "decode_ms %.3f" % d["kernel_share"]["decode_ms_per_step"], "layout_ms %.3f" % d["kernel_share"]["layout_ms_per_step"], "fallbacks", d["exact_fallbacks"]["count"])
PY
done
