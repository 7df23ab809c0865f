#!/bin/bash
# The ncu evidence of a round: launch list of timed steps + one --set full capture of every hot kernel, exported as CSV
# (the .ncu-rep files exceed what gpurun copies back).  Usage (under gpurun): bash tools/ncu_profiles.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --other none --no-cpu-baseline --no-host-state --settle 0"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_${tag}.csv $B --steps 4 --warmup 4 > /dev/null 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:"move_stream|cell_kernel|scatter_index|segment_sort|ntc_kernel|scan_final_self" -s 24 -c 6 -f -o gpurun_out/prof_${tag}_couette $B --steps 3 --warmup 3 > /dev/null 2>> gpurun_out/prof_${tag}.err
$N -k regex:"move_stream|ntc_kernel|cell_kernel" -s 12 -c 3 -f -o gpurun_out/prof_${tag}_n2lb3d $B --steps 3 --warmup 3 --case box --gas n2lb > /dev/null 2>> gpurun_out/prof_${tag}.err
$N -k regex:"bgk_kernel" -s 3 -c 1 -f -o gpurun_out/prof_${tag}_bgk3d $B --steps 3 --warmup 3 --case box --collision bgk > /dev/null 2>> gpurun_out/prof_${tag}.err
for r in gpurun_out/prof_${tag}_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  ncu -i $r --page source --csv --print-source cuda,sass > ${b}_src.csv 2>/dev/null
  rm -f $r
done
python tools/launch_summary.py gpurun_out/launches_${tag}.csv
ls -la gpurun_out | grep ${tag}
