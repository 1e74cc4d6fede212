#!/usr/bin/env bash
# round 2: ncu launch list + one-frame ncu --set full capture (with source) of the round-2 kernels at C3 full
set -u
O=gpurun_out/r2p; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_C3.csv \
    python bench.py --steps 2 --warmup 3 --cpu-frames 0 --no-extra --no-parity > $O/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_|preprocess_|scatter_entries|sort_tiles|scan_tiles' -s 24 -c 8 -f -o $O/prof_C3 \
    python bench.py --steps 2 --warmup 3 --cpu-frames 0 --no-stage-timing --no-extra --no-parity > $O/ncu_full_C3.log 2>&1; echo "ncu full rc=$?"
tail -3 $O/ncu_full_C3.log
ncu -i $O/prof_C3.ncu-rep --page raw --csv > $O/prof_C3_raw.csv 2>/dev/null
ncu -i $O/prof_C3.ncu-rep --page source --csv > $O/prof_C3_src.csv 2>/dev/null
ls -la $O
