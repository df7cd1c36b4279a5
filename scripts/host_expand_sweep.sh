#!/bin/bash
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled
cat /proc/sys/kernel/numa_balancing
for T in 8 16 32 64; do
  CW_UNPACK_THREADS=$T python scripts/host_expand_bench.py 192 2>&1 | tee -a gpurun_out/host_expand.log
done
CW_UNPACK_THREADS=32 CW_UNPACK_PIN=0 python scripts/host_expand_bench.py 192 2>&1 | tee -a gpurun_out/host_expand.log
CW_UNPACK_THREADS=16 CW_UNPACK_PIN=0 python scripts/host_expand_bench.py 192 2>&1 | tee -a gpurun_out/host_expand.log
CW_UNPACK_THREADS=32 CW_EXPAND_ISA=256 python scripts/host_expand_bench.py 192 2>&1 | tee -a gpurun_out/host_expand.log
