#!/bin/bash
# polyhedron neighbour-build sweep: variants in build/variants/ x grid cell sizes (restores the product .so afterwards)
cp exadem_b200/libexadem_b200.so /tmp/product.so
for v in build/variants/*.so; do
  cp $v exadem_b200/libexadem_b200.so
  for cs in 2.0 1.6 1.4; do
    echo -n "== $(basename $v .so) cell $cs: "
    python bench.py --workload poly --particles 500000 --cell-size $cs --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step %.4f rebuild_once_ms %.2f' % (d['ms_per_step'], r['rebuild_once_ms']), {k: round(v,4) for k,v in r['per_step_ms'].items() if isinstance(v, float)})"
  done
done
cp /tmp/product.so exadem_b200/libexadem_b200.so
