#!/bin/bash
# diagnostic: how fast is the tape interpreter when its slot traffic stays in L2 (CW_DEBUG_WRAP folds the
# slot index into a small window; results are garbage, only the time is of interest)
run() {
  echo "== $1"
  CW_BENCH_NOCHECK=1 python bench.py --steps 4 --warmup 3 --batch-per-gpu 1024 --no-cpu-baseline --e2e-steps 0 --no-r1cs 2>&1 | python scripts/show_bench.py
}
build() {
  CW_NVCC_EXTRA="$1" python -c "
import sys; sys.path.insert(0,'.')
from circom_b200 import build; build.build(force=True, verbose=True)" 2>&1 | grep -A2 "tape_exec_kernelILi0ELb0" | grep -E "registers" | tr '\n' ' '
  echo
}
build "-DCW_DEBUG_WRAP=0x7FF"
run "wrap 2048 slots (64 KB per instance, 64 MB total: L2-resident)"
build "-DCW_DEBUG_WRAP=0xFFFF"
run "wrap 65536 slots (2 MB per instance, 2 GB total: DRAM, dense)"
build ""
