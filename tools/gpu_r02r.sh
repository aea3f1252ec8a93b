#!/bin/bash
# r02r (1 GPU): A/B on one box: 1d level layout [p0 p1 p2] per cell (lib_old = HEAD) vs [(p1,p2)] [p0] (lib), alternating, kernel ms
mkdir -p gpurun_out
for rep in 1 2 3; do
for L in lib_old lib; do
for W in C1 C2; do
NUFI_B200_LIB=$PWD/numericalflowiteration_b200/$L/libnufi_b200.so timeout 300 python tools/sweep.py $W --reps 20 2>&1 | tail -1 | sed "s/^/$L $W: /"
done; done; done > gpurun_out/r02r_ab.txt
cat gpurun_out/r02r_ab.txt
