#!/bin/bash
# BGK kernel check: parity tests of the relaxation path, then the box BGK case and BASELINE config 4.  Usage (under gpurun): bash tools/ab_bgk.sh <tag>
tag=${1:-bgk}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "bgk or hybrid or macro or interpolation or tutorial or plate" > gpurun_out/${tag}_tests.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_tests.txt
tail -4 gpurun_out/${tag}_tests.txt
B="python bench.py --no-cpu-baseline --no-host-state --settle 0"
$B --other none --case box --collision bgk --steps 10 --warmup 5 > gpurun_out/${tag}_box.json 2> gpurun_out/${tag}.err
$B --other config4 --steps 5 --warmup 3 > gpurun_out/${tag}_c4.json 2>> gpurun_out/${tag}.err
python - <<PY
import json
for n in ("box","c4"):
    d=json.loads(open(f"gpurun_out/${tag}_{n}.json").read().strip().splitlines()[-1])
    print(n, d["ms_per_step"], d["roofline"]["phase_ms"])
    oc=d.get("other_configs") or {}
    for k,v in oc.items(): print(k, v.get("ms_per_step"), v.get("phase_ms"), v.get("error"))
PY
