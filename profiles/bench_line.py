"""Pretty-print the interesting fields of a bench.py JSON line read from stdin."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print("GTEPS %.1f  ms/step %.2f  ms/pass %.3f  frac %.3f  e2e %.1f  launches %s  bfs %s  cpu %s" % (
        d["value"], d["ms_per_step"], r.get("ms_per_launch", 0), r.get("frac", 0), d["e2e"]["value"],
        d.get("gpu_launches"), d.get("bfs", {}).get("gteps"), d.get("cpu_baseline", {}).get("value")))
