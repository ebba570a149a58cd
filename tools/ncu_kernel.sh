#!/bin/bash
# ncu --set full on one kernel of the bench (1024 proofs, one step).  usage: tools/ncu_kernel.sh <kernel-regex> [skip]
K=$1; S=${2:-2}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --proofs 1024 --no-cpu-baseline --no-secondary > gpurun_out/plain_$K.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/prof_$K \
    python bench.py --steps 1 --warmup 3 --proofs 1024 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$K.log 2>&1
tail -3 gpurun_out/ncu_$K.log
