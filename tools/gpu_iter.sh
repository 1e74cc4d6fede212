#!/usr/bin/env bash
# Quick iteration on the GPU box: parity tests + one bench line (+ optional ncu of the render kernels).
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_iter.sh [ncu]'
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 --cpu-frames 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --steps 30 --warmup 5 --cpu-frames 0 --variant light > gpurun_out/bench_light.json 2>> gpurun_out/bench.err; echo "bench light rc=$?"
cat gpurun_out/bench_light.json
if [ "${1:-}" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_ -s 6 -c 2 -o gpurun_out/prof_render \
      python bench.py --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
