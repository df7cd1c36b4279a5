#!/bin/bash
# forwarding ring on / off and CTA shapes (tape kernel only)
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 0"
run() { echo "== $1"; shift; env "$@" 2>&1 | python scripts/show_bench.py; }
run "ring on, batch 1024 (+r1cs check of the witnesses)" $B
run "ring off, batch 1024" CW_RING=0 $B --no-r1cs
run "ring on, batch 1184" $B --no-r1cs --batch-per-gpu 1184
run "ring on, batch 2048, 64 threads" CW_THREADS=64 $B --no-r1cs --batch-per-gpu 2048
run "ring on, batch 2048, 128 threads" CW_THREADS=128 $B --no-r1cs --batch-per-gpu 2048
