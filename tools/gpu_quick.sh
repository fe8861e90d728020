#!/bin/bash
# parity tests + a short steady-state bench (no e2e, no cpu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/bench_quick.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_quick.json'))
t, n = d["roofline"], d["roofline_other"]
if t["bound"] != "hbm": t, n = n, t
print("pos/s %.0f evals/launch %.0f net_ms %.4f tree_ms %.4f hit %.3f sims/launch %.0f" % (
    d["value"], n["evals_per_launch"], n["ms_per_launch"], t["ms_per_launch"], d["memo_hit_rate"], t["sims_per_launch"]))
PY
