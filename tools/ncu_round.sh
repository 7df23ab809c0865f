#!/bin/bash
# ncu --set full captures of the hot kernels on several shapes.  Usage (under gpurun): bash tools/ncu_round.sh <tag>
tag=${1:-ncu}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --settle 0 --other none --no-cpu-baseline --no-host-state"
N="ncu --set full --clock-control none --import-source on"
UGF_MOVE_V2=0 $N -k regex:"move_stream" -s 6 -c 1 -f -o gpurun_out/${tag}_couette_v1 $B > /dev/null 2>> gpurun_out/${tag}.err
UGF_MOVE_V2=1 $N -k regex:"move_stream" -s 6 -c 1 -f -o gpurun_out/${tag}_couette_v2 $B > /dev/null 2>> gpurun_out/${tag}.err
UGF_MOVE_V2=0 $N -k regex:"move_stream|ntc_kernel" -s 12 -c 2 -f -o gpurun_out/${tag}_n2lb_v1 $B --case box --gas n2lb > /dev/null 2>> gpurun_out/${tag}.err
UGF_MOVE_V2=0 $N -k regex:"bgk_kernel" -s 6 -c 1 -f -o gpurun_out/${tag}_bgk $B --case box --collision bgk > /dev/null 2>> gpurun_out/${tag}.err
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/${tag}.err
# the reports are too large to travel (64 MiB limit): export the pages read afterwards and drop the reports
for r in gpurun_out/${tag}_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  ncu -i $r --page source --csv --print-source cuda,sass > ${b}_src.csv 2>/dev/null
  rm -f $r
done
ls -la gpurun_out/ | head -30
