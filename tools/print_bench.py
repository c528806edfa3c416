"""One-line digest of a bench.py JSON line.  Usage: python tools/print_bench.py bench.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
r = d.get("roofline", {})
s = r.get("stage_ms_per_step", {})
print("value %.4g %s | %.3f ms/step | stages %s | truncated %s | e2e %.4g | roofline %s %.3f | cpu %s | launches %s" % (
    d["value"], d["unit"], d["ms_per_step"], dict((k, round(v, 2)) for k, v in s.items() if v > 0.05),
    (d.get("truncated_iteration") or {}).get("ms_per_step"), d["e2e"]["value"], r.get("bound"), r.get("frac", 0.0),
    (d.get("cpu_baseline") or {}).get("value"), d.get("gpu_launches")))
