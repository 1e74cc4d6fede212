#!/usr/bin/env bash
# ncu --set full capture of the per-Gaussian / binning kernels at C3 full (one launch each)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'preprocess_|scatter_entries|sort_tiles|scan_tiles' -s 15 -c 5 -f -o gpurun_out/prof_misc2 \
    python bench.py --steps 2 --warmup 3 --cpu-frames 0 --no-stage-timing > gpurun_out/ncu_misc2.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_misc2.log
ncu -i gpurun_out/prof_misc2.ncu-rep --page raw --csv > gpurun_out/prof_misc2_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_misc2.ncu-rep --page source --csv > gpurun_out/prof_misc2_src.csv 2>/dev/null
ls -la gpurun_out/prof_misc2*
