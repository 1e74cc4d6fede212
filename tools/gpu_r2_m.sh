#!/usr/bin/env bash
# round 2: accumulator clear forked in front of the forward blend kernel (early_acc_clear): tests + A/B
set -u
O=gpurun_out/r2m; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "accumulator or golden or speculative or tracking" > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt; tail -n 3 $O/pytest_gpu.txt
bash tools/gpu_ab_opts.sh "C3 full;C3 light;C4 full" "early_acc_clear=1" "early_acc_clear=0" 2>&1 | tee $O/ab_acc.txt
python tools/ab_small.py --opt early_acc_clear --values 0,1 --configs C2 --rounds 4 2>&1 | tee $O/ab_acc_small.txt
