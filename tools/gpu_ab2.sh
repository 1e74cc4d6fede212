#!/usr/bin/env bash
# parity suite + A/B of one library option at C3 full/light and C4 full
set -u
opt=${1:-bulk_sh}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
bash tools/gpu_ab.sh $opt "C3 full" "C3 light" "C2 light" "C4 full"
