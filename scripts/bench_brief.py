#!/usr/bin/env python
"""Run bench.py with the given extra args and print a one-line digest (tuning aid)."""
import json
import subprocess
import sys

out = subprocess.run([sys.executable, "bench.py", *sys.argv[1:]], capture_output=True, text=True)
lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
if not lines:
    print("bench failed:", out.stderr[-2000:])
    sys.exit(1)
l = json.loads(lines[-1])
print(" ".join(sys.argv[1:]), "| value", round(l["value"]), "SPF/s | e2e", round(l["e2e"]["value"]),
      "| kernel_ms", round(l["roofline"]["kernel_ms"], 3), "| frac", round(l["roofline"]["frac"], 4),
      "| clocks", l["clocks"]["sm_mhz"], l["clocks"]["reasons"])
