run() {
  g=$1; shift
  env "$@" python bench.py --games $g --steps 2 --warmup 3 --preroll 30000 --passes 2000 --no-e2e --no-cpu 2>&1 | tail -1 > /tmp/line.json
  python - "$g $*" <<'PY'
import sys, json
d = json.load(open('/tmp/line.json'))
t, n = d["roofline"], d["roofline_other"]
if t["bound"] != "hbm": t, n = n, t
print("%-52s pos/s %.0f evals/launch %.0f net_ms %.4f tree_ms %.4f hit %.3f" % (
    sys.argv[1], d["value"], n["evals_per_launch"], n["ms_per_launch"], t["ms_per_launch"], d["memo_hit_rate"]))
PY
}
run 2048 A=1
run 8192 A=1
run 16384 A=1
run 16384 C4_CYCLE_LIMIT=80000
run 32768 A=1
