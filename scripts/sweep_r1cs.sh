#!/bin/bash
# R1CS check variants (lane-group kernel for long rows on / off, instances per block)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0"
run() { echo "== $1"; shift; env "$@" 2>&1 | python scripts/show_bench.py; }
run "split on (default)" $B
run "split off" CW_R1CS_SPLIT=0 $B
run "split on, ipb 4" CW_R1CS_IPB=4 $B
run "split on, ipb 16" CW_R1CS_IPB=16 $B
