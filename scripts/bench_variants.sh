#!/bin/bash
# bench every .so in _variants/ against the main build (debug aid)
cp rascal_b200/librascal_b200.so /tmp/lib_main.so
for f in /tmp/lib_main.so _variants/*.so; do
  cp $f rascal_b200/librascal_b200.so; touch rascal_b200/librascal_b200.so
  echo "== $f"; timeout 300 python bench.py --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g ms/step %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step']))"
done
cp /tmp/lib_main.so rascal_b200/librascal_b200.so
