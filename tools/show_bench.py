#!/usr/bin/env python
"""Print the headline numbers and per-stage times of a bench.py JSON line:  python tools/show_bench.py FILE [label]"""
import json
import sys

d = json.load(open(sys.argv[1]))
label = sys.argv[2] if len(sys.argv) > 2 else ""
print(label, round(d["value"]), round(d["e2e"]["value"]), {k: round(v["ms"], 3) for k, v in d["stages"].items()
                                                            if k in ("polar2cart", "scan_to_l0l1", "pyr_down", "klt", "reject")})
