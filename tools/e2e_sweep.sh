#!/bin/bash
# e2e throughput of the host-pointer entry for several pipeline piece sizes
for p in "$@"; do
  SRSLTE_B200_PIECE=$p timeout 200 python bench.py --blocks 32768 --e2e-blocks 32768 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > /tmp/e2e.json
  python - "$p" <<'PY'
import json,sys
d=json.loads(open('/tmp/e2e.json').read().strip().splitlines()[-1])
print("piece", sys.argv[1], "e2e Gbit/s %.2f" % d["e2e"]["value"], "ms %.2f" % d["e2e"]["ms_per_step"])
PY
done
