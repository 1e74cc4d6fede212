#!/usr/bin/env bash
# (historical record of a round-2 GPU call: the A/B option it toggles was removed after the measurement, see profiles/r02_ab_*.txt)
# round 2: A/B of compile-time variants (option tune: bit 0 = sort 6 CTAs/SM, bit 1 = -full forward blend 7 CTAs/SM)
set -u
O=gpurun_out/r2o; mkdir -p $O
bash tools/gpu_ab_opts.sh "C3 full;C4 full" "tune=0" "tune=1" "tune=2" 2>&1 | tee $O/ab_tune.txt
