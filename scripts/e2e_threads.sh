#!/bin/bash
# end-to-end leg only, for several sizes of the host-side expansion pool (development; output in gpurun_out/)
mkdir -p gpurun_out
for T in ${@:-8 16 32 64}; do
  CW_UNPACK_THREADS=$T python bench.py --no-configs --no-cpu-baseline --no-r1cs --steps 1 --warmup 1 --e2e-batch 4096 2>gpurun_out/e2e_T$T.err | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('threads $T', 'e2e', round(j['e2e']['value']), 'w/s chunk', j['e2e']['chunk'], 'value', round(j['value']))
" | tee -a gpurun_out/e2e_threads.log
done
