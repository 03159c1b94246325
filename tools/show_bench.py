#!/usr/bin/env python3
"""One-line digest of a bench.py JSON line (or the tail of its stderr if it is not JSON)."""
import json
import sys

try:
    d = json.load(open(sys.argv[1]))
    print(d["config"]["workload"][:40], "| n_gpus", d["n_gpus"], "| value", d["value"], d["unit"], "| ms/step", d["ms_per_step"], d["scaling"],
          "| roofline", d["roofline"]["achieved"], d["roofline"]["unit"], d["roofline"]["frac"], "| e2e", d["e2e"]["value"],
          "| clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
    for k, v in (d.get("also") or {}).items():
        rl, e = v.get("roofline") or {}, v.get("e2e") or {}
        print("   also", k, "| value", v.get("value"), "| ms/step", v.get("ms_per_step"), "| roofline", rl.get("achieved"), rl.get("unit"), rl.get("frac"),
              "| e2e", e.get("value"), e.get("path"), e.get("frac_of_bound"), "| cpu", (v.get("cpu_baseline") or {}).get("value"), v.get("error") or "")
except Exception as e:  # noqa: BLE001
    print("NOT A BENCH LINE:", e)
    if len(sys.argv) > 2:
        print(open(sys.argv[2]).read()[-2000:])
