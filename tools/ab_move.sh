#!/bin/bash
# A/B of the streamed move kernels on the bench shapes.  Usage (under gpurun): bash tools/ab_move.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out
B="python bench.py --other none --no-cpu-baseline --no-host-state --settle 10"
for v in 0 1; do
  UGF_MOVE_V2=$v $B > gpurun_out/${tag}_couette_v$v.json 2>> gpurun_out/${tag}.err
  UGF_MOVE_V2=$v $B --case box > gpurun_out/${tag}_box_v$v.json 2>> gpurun_out/${tag}.err
  UGF_MOVE_V2=$v $B --case box --gas n2lb > gpurun_out/${tag}_n2lb_v$v.json 2>> gpurun_out/${tag}.err
  UGF_MOVE_V2=$v $B --case cylinder > gpurun_out/${tag}_cyl_v$v.json 2>> gpurun_out/${tag}.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_*_v*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]["phase_ms"]
        print(f.split("/")[-1], "value %.3f G  ms %.4f  move %.4f sort %.4f cell %.4f coll %.4f" % (d["value"]/1e9, d["ms_per_step"], r["move"], r["sort"], r["cell"], r["collide"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/${tag}.err
