#!/usr/bin/env bash
# round 2, call 1: time the cp.async prefetch variant, parity-check it, run the sanitizers
set -u
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
bash tools/gpu_ab.sh bwd_prefetch "C3 full" "C3 light" "C2 light" "C4 full" > $O/ab_prefetch.txt 2>&1
GSR_TEST_OPTS=bwd_prefetch=1,bwd_packed=1 timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "golden or oracle or live or packed" > $O/pytest_prefetch.txt 2>&1
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > $O/sanitize_$tool.txt 2>&1
  echo "exit $?" >> $O/sanitize_$tool.txt
done
tail -3 $O/ab_prefetch.txt $O/pytest_prefetch.txt $O/sanitize_*.txt
