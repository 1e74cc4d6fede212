#!/usr/bin/env bash
# One gpurun call: GPU tests, golden vectors from the reference build, both bench arms, ncu lists.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_first.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
python tests/golden/make_golden.py --out gpurun_out/golden > gpurun_out/make_golden.log 2>&1; echo "golden rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_ref.json
python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_ -s 6 -c 4 -o gpurun_out/prof_render \
    python bench.py --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out
