#!/bin/bash
# polyhedron neighbour-build sweep over the variants in build/variants/ (restores the product .so afterwards)
cp exadem_b200/libexadem_b200.so /tmp/product.so
for v in build/variants/*.so; do
  cp $v exadem_b200/libexadem_b200.so
  echo -n "== $(basename $v .so): "
  python bench.py --workload poly --particles 500000 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step %.4f rebuild_once_ms %.2f' % (d['ms_per_step'], r['rebuild_once_ms']))"
done
cp /tmp/product.so exadem_b200/libexadem_b200.so
