#!/usr/bin/env bash
# round 2: one-frame ncu --set full capture of the -light variant at C3
set -u
O=gpurun_out/r2lp; mkdir -p $O
timeout 900 ncu --set full --clock-control none -k regex:'render_|preprocess_' -s 16 -c 4 -f -o $O/prof_C3_light \
    python bench.py --variant light --steps 2 --warmup 3 --cpu-frames 0 --no-stage-timing --no-extra --no-parity > $O/ncu_full_C3_light.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_C3_light.ncu-rep --page raw --csv > $O/prof_C3_light_raw.csv 2>/dev/null
rm -f $O/prof_C3_light.ncu-rep
ls -la $O
