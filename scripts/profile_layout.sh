#!/bin/bash
# ncu captures of the interpreter and the R1CS check at one layout point (development; outputs in gpurun_out/)
# usage: scripts/profile_layout.sh "<compact,bt,threads,batch>" <tag>
P=${1:-"1,5,256,18944"}
TAG=${2:-bt5}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tape_exec -s 1 -c 1 -o gpurun_out/prof_tape_$TAG -f \
    python scripts/sweep_layout.py --points "$P" --steps 1 --no-r1cs --out gpurun_out/ncu_scratch.jsonl > gpurun_out/prof_tape_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:r1cs_check -s 1 -c 1 -o gpurun_out/prof_r1cs_$TAG -f \
    python scripts/sweep_layout.py --points "$P" --steps 1 --out gpurun_out/ncu_scratch.jsonl > gpurun_out/prof_r1cs_$TAG.log 2>&1
ls -la gpurun_out | tail -8
