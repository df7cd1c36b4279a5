#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe); outputs land in gpurun_out/
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tape_exec -s 1 -c 1 -o gpurun_out/prof_tape_exec -f $CMD > gpurun_out/prof_tape.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:r1cs_check_kernel -s 1 -c 1 -o gpurun_out/prof_r1cs -f $CMD > gpurun_out/prof_r1cs.log 2>&1
ncu --set full --clock-control none -k regex:r1cs_bool -s 1 -c 1 -o gpurun_out/prof_r1cs_bool -f $CMD > gpurun_out/prof_r1cs_bool.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/launches_bench.log
