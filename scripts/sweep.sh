#!/bin/bash
# runs the quick bench once per tuning variant in build/variants/ (restores the product .so afterwards)
cp exadem_b200/libexadem_b200.so /tmp/product.so
for v in build/variants/*.so; do
  cp $v exadem_b200/libexadem_b200.so
  echo "== $(basename $v .so)"
  python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-poly "$@" 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.4g  ms/step %.4f' % (d['value'], d['ms_per_step']), {k: round(v,4) for k,v in r['per_step_ms'].items() if isinstance(v, float)})"
done
cp /tmp/product.so exadem_b200/libexadem_b200.so
