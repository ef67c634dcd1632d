#!/bin/sh
# same box, several builds of the library: small resident searches and the 50-tile root-parallel instance
for rep in 1 2; do
for lib in "$@"; do
  export BP_LIB=$PWD/visapi_b200/librectpack_b200_$lib.so
  echo "== $lib (rep $rep)"
  timeout 120 python tools/auto_threshold.py 2>&1 | grep '"P": 1,\|"P": 64,' | tr '\n' ' '; echo
  timeout 120 python tools/bench_root_parallel.py --exchange resident 2>&1 | tail -1 | grep -o '"ms_per_instance": [0-9.]*' | tr '\n' ' '; echo
done
done
