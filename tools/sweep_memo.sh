#!/bin/bash
# env-var sweeps over the engine's tuning knobs (C4_BUDGET, C4_CYCLE_LIMIT, C4_MEMO_LOG2, C4_POOLS, C4_NET_CTAS)
run() {
  env "$@" python bench.py --steps 2 --warmup 3 --preroll 30000 --passes 2000 --no-e2e --no-cpu 2>&1 | tail -1 > /tmp/line.json
  python - "$*" <<'PY'
import sys, json
d = json.load(open('/tmp/line.json'))
t, n = d["roofline"], d["roofline_other"]
if t["bound"] != "hbm": t, n = n, t
print("%-52s pos/s %.0f evals/launch %.0f net_ms %.4f tree_ms %.4f hit %.3f" % (
    sys.argv[1], d["value"], n["evals_per_launch"], n["ms_per_launch"], t["ms_per_launch"], d["memo_hit_rate"]))
PY
}
for a in "$@"; do run $a; done
