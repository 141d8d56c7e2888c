#!/usr/bin/env python
"""Print the essentials of bench.py JSON lines read from stdin."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    c = d.get("config", {})
    r = d.get("roofline", {})
    print("%s | %s | value %.3e | incl.eval %.3e | ms/step %.1f | e2e %s | roof %.3f | clk %s %s" % (
        c.get("workload", "?")[:3], c.get("table_store", d.get("impl", "?")), d["value"], c.get("env_steps_per_s_incl_eval", 0), d["ms_per_step"],
        ("%.3e" % d["e2e"]["value"]) if "e2e" in d else "-", r.get("frac", 0), d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons")))
