#!/bin/bash
# A/B of two builds of the library on the same box: tools/ab.sh libA.so libB.so  (alternating runs of the headline shapes)
for i in 1 2; do
  for lib in "$@"; do
    echo "== $lib"
    B2PIV_LIB=$PWD/$lib python tools/quick_bench.py --configs 2>&1 | sed -n '2,3p;5,6p' | sed 's/variant 0 run_len 0  //'
  done
done
