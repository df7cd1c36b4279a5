#!/bin/bash
# ncu capture of the interpreter on the function-hint variant of the headline circuit (run under gpurun)
mkdir -p gpurun_out
T="timeout 400"
$T ncu --set full --clock-control none --import-source on -k regex:tape_exec -s 1 -c 1 -o gpurun_out/r02_prof_tape_calls -f \
  python scripts/sweep_layout.py --workload ecdsa_scale_calls --points "1,5,0,18944" --steps 1 --no-r1cs --out gpurun_out/prof_calls_sweep.jsonl > gpurun_out/r02_prof_tape_calls.log 2>&1
f=gpurun_out/r02_prof_tape_calls.ncu-rep
ncu -i $f --page details > gpurun_out/r02_prof_tape_calls_details.txt 2>/dev/null
ncu -i $f --page source --csv > gpurun_out/r02_prof_tape_calls_source.csv 2>/dev/null
rm -f $f
ls -la gpurun_out
