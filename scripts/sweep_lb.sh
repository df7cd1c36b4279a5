#!/bin/bash
# register-budget sweep of the tape interpreter: __launch_bounds__(128, MB) caps registers so MB CTAs fit per SM
for mb in 8 10 12 14; do
  CW_NVCC_EXTRA="-DCW_TAPE_LB=128,$mb" python -c "
import sys; sys.path.insert(0,'.')
from circom_b200 import build; build.build(force=True, verbose=True)" 2>&1 | grep -A2 "tape_exec_kernelILi0ELb0" | grep -E "spill|registers" | tr '\n' ' '
  echo
  for B in 1536 2048; do
    echo "== minblocks $mb batch $B threads 128"
    CW_THREADS=128 python bench.py --steps 3 --warmup 2 --batch-per-gpu $B --no-cpu-baseline --e2e-steps 0 --no-r1cs 2>&1 | python scripts/show_bench.py
  done
done
