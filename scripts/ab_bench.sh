#!/bin/bash
# Development aid: bench.py with the current library and with another build (CAGC_LIB), alternating, on one box.
#   scripts/ab_bench.sh <other.so> [rounds]
other=$1; rounds=${2:-2}
for r in $(seq $rounds); do
  for tag in cur other; do
    if [ $tag = other ]; then export CAGC_LIB=$other; else unset CAGC_LIB; fi
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu --no-second-mode --sustained-steps 0 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
  done
done
