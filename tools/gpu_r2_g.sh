#!/usr/bin/env bash
# round 2, call 7: octet backward kernel (bwd_packed=4): parity subset + A/B against the quarter-list kernel
set -u
O=gpurun_out/r2g; mkdir -p $O
GSR_TEST_OPTS=bwd_packed=4 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or live or oracle or c_abi or linear" > $O/pytest_octet.txt 2>&1; echo "pytest exit $?" >> $O/pytest_octet.txt
tail -n 15 $O/pytest_octet.txt
bash tools/gpu_ab_opts.sh "C3 full;C3 light;C2 light;C4 full" "bwd_packed=2" "bwd_packed=4" 2>&1 | tee $O/ab.txt
