#!/usr/bin/env python
"""Print the per-stage table of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"ms/step {d['ms_per_step']:.3f}   e2e {(d["e2e"].get("ms_per_step") or float("nan")):.3f}   launches {d.get('gpu_launches')}  "
      f"roofline {d['roofline']['kernel']} {d['roofline']['frac']:.3f}")
tot = 0.0
for k, v in d["stages"].items():
    tot += v["ms_per_step"]
    print(f"{k:24s} {v['ms_per_step']:7.3f} ms  {v['GBps'] or 0:7.0f} GB/s  ({v['alg_GB_per_step']:.3f} GB)")
print(f"{'sum of stages':24s} {tot:7.3f} ms")
