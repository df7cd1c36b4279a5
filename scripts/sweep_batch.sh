nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader,nounits 2>&1 | head -3
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.hw_thermal_slowdown,clocks_throttle_reasons.sw_thermal_slowdown,clocks_throttle_reasons.sw_power_cap --format=csv,noheader,nounits 2>&1 | head -3
nvidia-smi --help-query-gpu 2>/dev/null | grep -i -E "reasons\.(active|hw_slowdown|sw_power)" | head
for cfg in "64 256" "64 512" "64 1024" "128 512" "128 1024" "256 512" "256 1024"; do
  set -- $cfg
  echo "== batch $1 threads $2"
  CW_THREADS=$2 python bench.py --steps 3 --warmup 1 --batch-per-gpu $1 --no-cpu-baseline --no-r1cs 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('wit/s %.0f ms %.2f exec %.2f gather %.2f e2e %.0f'%(j['value'],j['ms_per_step'],j['kernel_ms']['tape_exec+stage'],j['kernel_ms']['witness_gather'],j['e2e']['value']))
    else: print(l.strip()[-300:])
"
done
