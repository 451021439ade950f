#!/usr/bin/env python
"""One-line digest of bench JSON lines: python profiles/show.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    r = d.get("roofline") or {}; w = d.get("walk") or {}; a = d.get("alt_algo"); e2e = d.get("e2e") or {}
    print("%-44s N=%d %s value=%.3e step=%.3fms train_only=%.3fms kern=%.3fms frac=%.2f e2e=%.3e | walk %.3e steps/s %.3fms"
          % (f.split("/")[-1], d["n_gpus"], d.get("algo", "?"), d["value"], d["ms_per_step"],
             (d.get("train_only") or {}).get("ms_per_step", 0), r.get("kernel_ms", 0), r.get("frac", 0),
             e2e.get("value", 0) or 0, w.get("value", 0), w.get("ms_per_pass", 0))
          + (" | alt %s %.3e %.3fms" % (a["algo"], a["value"], a["ms_per_step"]) if a else ""))
