#!/usr/bin/env python
"""Print the headline fields of a bench.py JSON line read from stdin (tuning helper)."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(tag, d.get("n_gpus"), "ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"].get("ms_per_step", 0), 3),
          {k: round(v, 4) for k, v in d.get("kernel_ms_per_step", {}).items()}, "frame_frac", round(d.get("frame_roofline", {}).get("frac", 0), 4))
