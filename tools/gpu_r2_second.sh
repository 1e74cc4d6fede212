#!/usr/bin/env bash
# round 2, call 2: full GPU test suite, default bench (parity + extra configs), reference arm, initcheck, microbench
set -u
O=gpurun_out/r2b; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
timeout 900 python bench.py > $O/bench_b200.json 2> $O/bench_b200.err; echo "bench exit $?" >> $O/bench_b200.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
./tools/microbench/fp32x2 > $O/fp32x2.txt 2>&1
for opt in "bulk_sh=0" "bulk_sh=1"; do
  GSR_SANITIZE_SMALL=1 GSR_TEST_OPTS=$opt timeout 500 compute-sanitizer --tool initcheck --print-limit 5 python tools/sanitize_small.py > $O/initcheck_$opt.txt 2>&1
  echo "exit $?" >> $O/initcheck_$opt.txt
done
tail -n 4 $O/pytest_gpu.txt; tail -n 3 $O/bench_b200.err; cat $O/fp32x2.txt
