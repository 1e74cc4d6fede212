#!/usr/bin/env bash
# Round 2: one gpurun call collecting the evidence under gpurun_out/final2: smoke, GPU tests, both bench arms,
# ncu launch lists and full captures (C3 + C4), tracking benchmark, the other configs, sanitizers.
# Usage: gpurun --timeout 2400 -- 'bash tools/gpu_r2_final.sh'    then    python tools/collect_profiles.py
set -u
O=gpurun_out/final2
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python bench.py --steps 30 --warmup 5 --variant light --cpu-frames 0 --no-extra > $O/bench_light.json 2>> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_C3.csv \
    python bench.py --steps 2 --warmup 3 --cpu-frames 0 --no-extra --no-parity > $O/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
# one whole frame: memset, preprocess_fwd, scan, scatter, sort, render_fwd, (loss kernels are torch's), render_bwd, preprocess_bwd
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_|preprocess_|scatter_entries|sort_tiles|scan_tiles' -s 28 -c 7 -f -o $O/prof_C3 \
    python bench.py --steps 2 --warmup 3 --cpu-frames 0 --no-stage-timing --no-extra --no-parity > $O/ncu_full_C3.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/prof_C3.ncu-rep --page raw --csv > $O/prof_C3_raw.csv 2>/dev/null
ncu -i $O/prof_C3.ncu-rep --page source --csv > $O/prof_C3_src.csv 2>/dev/null
timeout 600 python tools/bench_tracking.py --config C2 --iters 50 > $O/tracking_C2.jsonl 2> $O/tracking.err; echo "tracking rc=$?"
timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file $O/tracking_launches.csv python tools/bench_tracking.py --arms tracker --iters 6 --reps 1 > $O/tracking_ncu.log 2>&1
bash tools/gpu_configs.sh > $O/configs.txt 2>&1; cp -r gpurun_out/configs $O/ 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_C4.csv \
    python bench.py --config C4 --steps 2 --warmup 3 --cpu-frames 0 --no-extra --no-parity > $O/bench_C4_under_ncu.log 2>&1; echo "ncu C4 list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'render_|preprocess_' -s 16 -c 4 -f -o $O/prof_C4 \
    python bench.py --config C4 --steps 2 --warmup 3 --cpu-frames 0 --no-stage-timing --no-extra --no-parity > $O/ncu_full_C4.log 2>&1; echo "ncu C4 full rc=$?"
ncu -i $O/prof_C4.ncu-rep --page raw --csv > $O/prof_C4_raw.csv 2>/dev/null
rm -f $O/prof_C4.ncu-rep
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > $O/sanitize_$tool.txt 2>&1
  echo "exit $?" >> $O/sanitize_$tool.txt
done
GSR_SANITIZE_SMALL=1 GSR_TEST_OPTS=bulk_sh=0 timeout 600 compute-sanitizer --tool initcheck --print-limit 5 python tools/sanitize_small.py > $O/sanitize_initcheck.txt 2>&1
echo "exit $?" >> $O/sanitize_initcheck.txt
for seed in 2 3; do
  timeout 900 python tools/parity_fuzz.py --cases 1500 --seed $seed --out $O/parity_fuzz_1500_seed$seed.txt > /dev/null 2>>$O/parity_fuzz.err; echo "fuzz seed $seed exit $?"
  tail -n 1 $O/parity_fuzz_1500_seed$seed.txt | cut -c1-300
done
tail -n 4 $O/sanitize_*.txt
cat $O/configs.txt | tail -8
ls -la $O | head -50
