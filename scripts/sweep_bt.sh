#!/bin/bash
# instance-tile width sweep (CW_BT_LOG2) at fixed batch
B=${B:-1024}
CFGS=${CFGS:-"0:128 1:128 1:256 2:128 2:256 3:256 3:512"}
for cfg in $CFGS; do
  BT=${cfg%%:*}; T=${cfg##*:}
  echo "== batch $B bt_log2 $BT threads $T"
  CW_BT_LOG2=$BT CW_THREADS=$T python bench.py --steps 3 --warmup 2 --batch-per-gpu $B --no-cpu-baseline --e2e-steps 0 --no-r1cs 2>&1 | python scripts/show_bench.py
done
