#!/bin/bash
# half-pool / CTA-cap sweep
run() {
  env "$@" python bench.py --steps 2 --warmup 3 --preroll 30000 --passes 2000 --no-e2e --no-cpu 2>&1 | tail -1 > /tmp/line.json
  python - "$*" <<'PY'
import sys, json
d = json.load(open('/tmp/line.json')); r = d["roofline"]
print("%-44s pos/s %.0f evals/launch %.0f net_ms %.4f tree_ms %.4f net TF %.0f" % (sys.argv[1], d["value"], r["evals_per_launch"], r["net_ms_per_launch"], r["tree_ms_per_launch"], r["achieved"]))
PY
}
run C4_POOLS=1
run C4_POOLS=2 C4_NET_CTAS=148
run C4_POOLS=2 C4_NET_CTAS=128
run C4_POOLS=2 C4_NET_CTAS=112
run C4_POOLS=2 C4_NET_CTAS=148 C4_BUDGET=3
run C4_POOLS=1 C4_BUDGET=1
