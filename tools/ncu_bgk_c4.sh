#!/bin/bash
# ncu --set full of the relaxation kernels on BASELINE config 4 (hybrid cylinder) at reduced scale.  Usage (under gpurun): bash tools/ncu_bgk_c4.sh <tag>
tag=${1:-r2c4}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-host-state --settle 0 --other config4 --other-scale 0.5 --steps 2 --warmup 1"
ncu --set full --clock-control none --import-source on -k regex:"bgk_" -s 4 -c 2 -f -o gpurun_out/prof_${tag} $B > /dev/null 2> gpurun_out/prof_${tag}.err
r=gpurun_out/prof_${tag}.ncu-rep
ncu -i $r --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i $r --page source --csv --print-source cuda,sass > gpurun_out/prof_${tag}_src.csv 2>/dev/null
rm -f $r
ls -la gpurun_out | grep ${tag}
