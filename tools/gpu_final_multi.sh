#!/bin/bash
# Final multi-GPU evidence: 2-GPU tests (N >= 2), the bench line at N ranks with the other BASELINE configs, optionally the reference arm.
# Usage (under gpurun --gpus N): bash tools/gpu_final_multi.sh <tag> <N> [ref]
tag=${1:-multi}; N=${2:-2}; ref=${3:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/${tag}_tests.txt 2>&1; tail -3 gpurun_out/${tag}_tests.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1500 $T --master-port 29533 bench.py --gpus $N > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/${tag}_bench.err
if [ -n "$ref" ]; then
  timeout 900 $T --master-port 29544 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err
  echo "ref exit $?"; tail -c 400 gpurun_out/${tag}_ref.err; tail -c 900 gpurun_out/${tag}_ref.json
fi
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench.json") if l.startswith("{")][-1])
    print("value %.3f G ms %.4f e2e %.3f G" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9), d.get("migration",{}).get("rounds_per_step"), d["checks"]["parcel_balance_ok"])
    for k,v in d["other_configs"].items():
        print(k, {kk: v.get(kk) for kk in ("value","ms_per_step","parcels","rounds_per_step","error","skipped")}, v.get("per_step"))
except Exception as e:
    print("ERR", e)
PY
