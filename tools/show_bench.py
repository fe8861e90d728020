"""print the interesting fields of a bench.py JSON line"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f  engine %s  ms/step %.1f  hit %.3f  launches %d" % (d["value"], d["config"]["engine"], d["ms_per_step"], d["memo_hit_rate"], d["gpu_launches"]))
if d.get("e2e"): print("e2e %.0f  (%.3f s, %d records)" % (d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"]["records"]))
for r in ("roofline", "roofline_other"):
    print(r, {k: d[r].get(k) for k in ("kernel", "achieved", "frac", "ms_per_launch", "share_of_step")})
for k in ("generation_1200", "steady_state", "engine_ab", "cpu_baseline"):
    if d.get(k): print(k, {a: b for a, b in d[k].items() if a in ("value", "seconds", "records_sha256_16", "engine", "memo_hit_rate", "cores", "n_gpus")})
