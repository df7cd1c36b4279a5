#!/bin/bash
# batch / CTA-size sweep of the bench workload (device-resident throughput)
for cfg in "256 512" "512 256" "1024 128" "1024 64" "2048 64" "2048 128" "3072 64"; do
  set -- $cfg
  echo "== batch $1 threads $2"
  CW_THREADS=$2 python bench.py --steps 3 --warmup 2 --batch-per-gpu $1 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('wit/s %.0f ms %.2f exec %.2f e2e %.0f r1cs %.0f Mc/s (%.2f ms) roofline %.3f'%(j['value'],j['ms_per_step'],j["kernel_ms"]["tape_exec+stage"],j['e2e']['value'],j['r1cs']['mconstraints_per_s'],j['r1cs']['ms'],j['roofline']['frac']))
    else: print(l.strip()[-300:])
"
done
