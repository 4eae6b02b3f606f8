#!/bin/bash
# e2e throughput of pm_engine_scan_host for several pipeline chunk sizes (development aid)
for m in 8 16 32 64; do
  PM_HOST_CHUNK_MIB=$m python bench.py --steps 3 --no-cpu-baseline 2>/dev/null > /tmp/e2e_$m.json
  python - "$m" <<'PY'
import sys, json
m = sys.argv[1]
d = json.loads(open(f"/tmp/e2e_{m}.json").read().strip().splitlines()[-1])
print("chunk MiB", m, "e2e", round(d["e2e"]["value"], 2), "value", round(d["value"], 1), d["record_gather"])
PY
done
