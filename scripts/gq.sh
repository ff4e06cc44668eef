#!/bin/bash
# quick GPU check: parity tests + short bench summary (used during kernel iteration)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e "$@" 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.4g  ms/step %.4f  frac %.3f  pp_ms %.4f' % (d['value'], d['ms_per_step'], r['frac'], r['mean_launch_ms'])); print(r['per_step_ms'])"
