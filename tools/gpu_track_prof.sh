#!/usr/bin/env bash
# Full GPU suite + tracker launch list (ncu, graph nodes profiled individually).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_tracking.py --config C2 --iters 50 > gpurun_out/bench_tracking.jsonl 2> gpurun_out/bench_tracking.err; echo "bench rc=$?"
cat gpurun_out/bench_tracking.jsonl
timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file gpurun_out/track_launches.csv python tools/bench_tracking.py --arms tracker --iters 6 --reps 1 \
  > gpurun_out/track_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/track_ncu.log
