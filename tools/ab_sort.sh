#!/bin/bash
tag=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" > gpurun_out/${tag}_tests.txt 2>&1; tail -6 gpurun_out/${tag}_tests.txt
B="python bench.py --other none --no-cpu-baseline --no-host-state --settle 10"
UGF_SORT_ORDERED=0 $B > gpurun_out/${tag}_couette_a_generic.json 2>> gpurun_out/${tag}.err
UGF_SORT_ORDERED=1 $B > gpurun_out/${tag}_couette_b_ordered.json 2>> gpurun_out/${tag}.err
UGF_SORT_ORDERED=0 $B --case box --gas n2lb > gpurun_out/${tag}_n2lb_a_generic.json 2>> gpurun_out/${tag}.err
UGF_SORT_ORDERED=1 $B --case box --gas n2lb > gpurun_out/${tag}_n2lb_b_ordered.json 2>> gpurun_out/${tag}.err
UGF_SORT_ORDERED=0 $B --case cylinder --weighted > gpurun_out/${tag}_cylw_a_generic.json 2>> gpurun_out/${tag}.err
UGF_SORT_ORDERED=1 $B --case cylinder --weighted > gpurun_out/${tag}_cylw_b_ordered.json 2>> gpurun_out/${tag}.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_*_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]["phase_ms"]
        print(f.split("/")[-1], "value %.3f G  ms %.4f  move %.4f sort %.4f cell %.4f coll %.4f" % (d["value"]/1e9, d["ms_per_step"], r["move"], r["sort"], r["cell"], r["collide"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 30 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 4 --settle 0 --other none --no-cpu-baseline --no-host-state > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv 2>/dev/null | head -20
