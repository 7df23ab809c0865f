#!/bin/bash
# A/B of tuning switches on the bench workload; prints value, ms/step and the per-phase device times.
run() {
  echo "== $*"
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-host-state $BENCH_ARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms', {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3))
    elif l: print(l)
"
}
for cfg in "$@"; do run $cfg; done
