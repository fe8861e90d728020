#!/bin/bash
run() {
  env "$@" python bench.py --steps 2 --warmup 3 --preroll 30000 --passes 2000 --no-e2e --no-cpu 2>&1 | tail -1 > /tmp/line.json
  python - "$*" <<'PY'
import sys, json
d = json.load(open('/tmp/line.json')); r = d["roofline"]
print("%-52s pos/s %.0f evals/launch %.0f net_ms %.4f tree_ms %.4f" % (sys.argv[1], d["value"], r["evals_per_launch"], r["net_ms_per_launch"], r["tree_ms_per_launch"]))
PY
}
for a in "$@"; do run $a; done
