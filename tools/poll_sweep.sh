#!/bin/sh
# resident kernel: poll back-off settings (BP_RESIDENT_POLL_NS / BP_RESIDENT_POLL_TIER) on small and large searches
for cfg in "256 0" "1024 0" "4096 0" "256 256" "256 512" "512 512" "256 1024"; do
  set -- $cfg
  export BP_RESIDENT_POLL_NS=$1 BP_RESIDENT_POLL_TIER=$2
  echo "== poll_ns=$1 tier=$2"
  timeout 120 python tools/auto_threshold.py 2>&1 | grep '"P": 1,\|"P": 64,\|"P": 250,' | sed 's/"stepped": [0-9.]*, //'
  timeout 120 python tools/shard_sweep.py --mode resident --reps 6 2>&1 | cut -c1-110
  timeout 120 python tools/bench_root_parallel.py --exchange resident 2>&1 | tail -1 | grep -o '"ms_per_instance": [0-9.]*' | tr '\n' ' '; echo
done
