#!/bin/bash
tag=${1:-ncus}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --settle 0 --other none --no-cpu-baseline --no-host-state"
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 32 --csv --log-file gpurun_out/${tag}_launches.csv $B > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv | head -20
ncu --set full --clock-control none --import-source on -k regex:"ordered_scatter|arrivals_sort" -s 4 -c 2 -f -o gpurun_out/${tag}_sort $B > /dev/null 2>> gpurun_out/${tag}.err
for r in gpurun_out/${tag}_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  ncu -i $r --page source --csv --print-source cuda,sass > ${b}_src.csv 2>/dev/null
  rm -f $r
done
