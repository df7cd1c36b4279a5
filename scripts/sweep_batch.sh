#!/bin/bash
# batch / CTA-size sweep of the bench workload (device-resident throughput)
CFGS=${CFGS:-"256:512 512:256 1024:128 1024:64 2048:64 2048:128 3072:64"}
for cfg in $CFGS; do
  B=${cfg%%:*}; T=${cfg##*:}
  echo "== batch $B threads $T"
  CW_THREADS=$T python bench.py --steps 3 --warmup 2 --batch-per-gpu $B --no-cpu-baseline --e2e-steps 0 2>&1 | python scripts/show_bench.py
done
