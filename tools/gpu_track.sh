#!/usr/bin/env bash
# Tracker tests + tracking benchmark on the GPU box.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tracking_gpu.py -m gpu -q -x > gpurun_out/pytest_track.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_track.log
timeout 600 python tools/bench_tracking.py --config C2 --iters 50 > gpurun_out/bench_tracking.jsonl 2> gpurun_out/bench_tracking.err; echo "bench rc=$?"
cat gpurun_out/bench_tracking.jsonl; tail -5 gpurun_out/bench_tracking.err
