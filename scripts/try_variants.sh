#!/bin/bash
# run one pytest selection against every .so in _variants/ (debug aid)
SEL=${1:-"tests/test_gpu_batch.py -k quality"}
cp rascal_b200/librascal_b200.so /tmp/lib_main.so
for f in /tmp/lib_main.so _variants/*.so; do
  cp $f rascal_b200/librascal_b200.so; touch rascal_b200/librascal_b200.so
  echo "== $f"; RB2_DEBUG_SYNC=1 timeout 300 python -m pytest $SEL -x -q -p no:cacheprovider 2>&1 | grep -i "rb2:\|passed\|failed" | head -5
done
cp /tmp/lib_main.so rascal_b200/librascal_b200.so
