#!/bin/bash
# headline bench on the other BASELINE widths (no secondary legs)
for w in cfg3-2048 cfg4; do
  timeout 200 python bench.py --workload $w --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', round(d['value'],1), round(d['ms_per_step'],2), round(d['step_roofline']['frac'],3), round(d['roofline']['achieved']))"
done
