#!/bin/bash
# One GPU session: tests, bench line, launch list.  Usage (under gpurun): bash tools/gpu_round.sh <tag> [pytest-args]
tag=${1:-run}
shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt; free -g | head -2 >> gpurun_out/${tag}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${tag}_tests.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_tests.txt
tail -5 gpurun_out/${tag}_tests.txt
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"
tail -c 3000 gpurun_out/${tag}_bench.err
