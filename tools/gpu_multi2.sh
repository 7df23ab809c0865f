#!/bin/bash
tag=${1:-m2}; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/${tag}_tests.txt 2>&1; tail -3 gpurun_out/${tag}_tests.txt
for f in 0 1; do
UGF_MIG_FUSED=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$f bench.py --gpus $N --other none > gpurun_out/${tag}_bench_fused$f.json 2> gpurun_out/${tag}_bench$f.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench_fused$f.json") if l.startswith("{")][-1])
    print("fused=$f value %.3f G ms %.4f e2e %.3f G launches %d" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["gpu_launches"]), d["checks"]["parcel_balance_ok"], d["migration"]["rounds_per_step"])
except Exception as e:
    print("ERR", e)
PY
done
python bench.py --other none --no-cpu-baseline --no-host-state --settle 10 > gpurun_out/${tag}_single.json 2>> gpurun_out/${tag}.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_single.json')); print('single', d['value']/1e9, d['roofline']['phase_ms'])"
tail -3 gpurun_out/${tag}_bench1.err
