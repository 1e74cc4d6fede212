#!/usr/bin/env bash
# round 2, call 3: quarter-list backward kernel: tests, A/B against the 8x8 packed kernel, 64- vs 72-register builds, ncu
set -u
O=gpurun_out/r2c; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -n 5 $O/pytest_gpu.txt
bash tools/gpu_ab_opts.sh "C3 full;C3 light;C2 light;C4 full;C4 light;C1 light" "bwd_packed=1" "bwd_packed=3,bwd_occ=8" "bwd_packed=3,bwd_occ=7" "bwd_packed=0" 2>&1 | tee $O/ab_bwdq.txt
# launch list + full capture of the two blend kernels at C3 full (default options)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_C3_full.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --no-stage-timing --no-extra --no-parity > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_' -s 6 -c 2 -f -o $O/prof_render python bench.py --steps 2 --warmup 3 --cpu-frames 0 --no-stage-timing --no-extra --no-parity > $O/ncu_full.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_render.ncu-rep --page raw --csv > $O/prof_render_raw.csv 2>/dev/null
ncu -i $O/prof_render.ncu-rep --page source --csv > $O/prof_render_src.csv 2>/dev/null
python tools/ncu_summary.py $O/prof_render_raw.csv $O/prof_render_src.csv > $O/ncu_summary_render.txt 2>&1
rm -f $O/prof_render.ncu-rep
head -60 $O/ncu_summary_render.txt
