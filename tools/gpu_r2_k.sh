#!/usr/bin/env bash
# (historical record of a round-2 GPU call: the A/B option it toggles was removed after the measurement, see profiles/r02_ab_*.txt)
# round 2, call 11: tiles handed out longest list first (tile_lpt): full GPU suite + A/B
set -u
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt; tail -n 5 $O/pytest_gpu.txt
bash tools/gpu_ab_opts.sh "C3 full;C3 light;C4 full" "tile_lpt=1" "tile_lpt=0" 2>&1 | tee $O/ab_lpt.txt
python tools/ab_small.py --opt tile_lpt --values 0,1 --configs C1,C2 --rounds 4 2>&1 | tee $O/ab_lpt_small.txt
python tools/bench_tracking.py --config C2 --iters 50 2>/dev/null | tail -n 3 | cut -c1-400
