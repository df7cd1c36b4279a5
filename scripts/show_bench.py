#!/usr/bin/env python
"""one-line summary of bench.py's JSON line (stdin)"""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        print(line.strip()[-300:])
        continue
    j = json.loads(line)
    out = "wit/s %.0f  ms/step %.2f  exec %.2f ms  e2e %.0f wit/s  roofline %.3f" % (
        j["value"], j["ms_per_step"], j["kernel_ms"]["tape_exec+stage"], j["e2e"]["value"] or 0, j["roofline"]["frac"])
    if "r1cs" in j:
        out += "  r1cs %.0f Mc/s (%.2f ms, frac %.3f)" % (j["r1cs"]["mconstraints_per_s"], j["r1cs"]["ms"],
                                                          j["r1cs"]["roofline"]["frac"])
    if "cpu_baseline" in j:
        out += "  cpu %.1f wit/s (%s, %d cores)" % (j["cpu_baseline"]["value"], j["cpu_baseline"]["kind"],
                                                    j["cpu_baseline"]["cores"])
    print(out)
