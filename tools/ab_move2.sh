#!/bin/bash
tag=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" > gpurun_out/${tag}_tests.txt 2>&1; tail -4 gpurun_out/${tag}_tests.txt
B="python bench.py --other none --no-cpu-baseline --no-host-state --settle 10"
run() { # name, env..., args
  name=$1; shift
  env "$@" > /dev/null 2>&1
}
UGF_MOVE_V2=0 $B > gpurun_out/${tag}_couette_a_v1.json 2>> gpurun_out/${tag}.err
UGF_MOVE_V2=1 $B > gpurun_out/${tag}_couette_b_v3.json 2>> gpurun_out/${tag}.err
UGF_MOVE_V2=1 UGF_MOVE_BPS=3 $B > gpurun_out/${tag}_couette_c_v3bps3.json 2>> gpurun_out/${tag}.err
UGF_MOVE_V2=1 $B > gpurun_out/${tag}_couette_d_v3again.json 2>> gpurun_out/${tag}.err
UGF_MOVE_V2=1 UGF_MOVE_BPS=2 $B --case box --gas n2lb > gpurun_out/${tag}_n2lb_c_v3bps2.json 2>> gpurun_out/${tag}.err
UGF_MOVE_V2=1 $B --case box --gas n2lb > gpurun_out/${tag}_n2lb_b_v3.json 2>> gpurun_out/${tag}.err
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
