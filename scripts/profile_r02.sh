#!/bin/bash
# ncu evidence of round 2 (B200_PROFILING.md recipe); run under gpurun, outputs land in gpurun_out/.
# The profiled command is the bench command with fewer steps; numbers printed under ncu are not bench values.
mkdir -p gpurun_out
T="timeout 280"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --no-configs --no-gather"
# every launch with its device time (headline workload)
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_launches_bench.log 2>&1
# the interpreter and the R1CS check, full metric set with source correlation
$T ncu --set full --clock-control none --import-source on -k regex:tape_exec -s 1 -c 1 -o gpurun_out/r02_prof_tape_exec -f $CMD > gpurun_out/r02_prof_tape.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:r1cs_check -s 1 -c 1 -o gpurun_out/r02_prof_r1cs -f $CMD > gpurun_out/r02_prof_r1cs.log 2>&1
# config C4 (Sha256(512), BLS12-381): both kernels
C4="python bench.py --workload sha256_512_bls --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0"
$T ncu --set full --clock-control none --import-source on -k regex:r1cs_check -s 1 -c 1 -o gpurun_out/r02_prof_r1cs_c4 -f $C4 > gpurun_out/r02_prof_r1cs_c4.log 2>&1
$T ncu --set full --clock-control none -k regex:tape_exec -s 1 -c 1 -o gpurun_out/r02_prof_tape_c4 -f $C4 > gpurun_out/r02_prof_tape_c4.log 2>&1
# the packed-transfer kernel (end-to-end leg on a small slice)
$T ncu --set full --clock-control none -k regex:witness_pack -s 2 -c 1 -o gpurun_out/r02_prof_pack -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-r1cs --e2e-steps 1 --e2e-batch 1024 --no-configs --no-gather > gpurun_out/r02_prof_pack.log 2>&1
# the reports themselves are ~20 MB each (gpurun brings back 64 MiB): export what is read, drop the reports
for f in gpurun_out/r02_prof_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page details > ${b}_details.txt 2>/dev/null
  ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null
  case $b in *tape_exec|*prof_r1cs|*r1cs_c4) ncu -i $f --page source --csv > ${b}_source.csv 2>/dev/null;; esac
  rm -f $f
done
ls -la gpurun_out | grep r02_
