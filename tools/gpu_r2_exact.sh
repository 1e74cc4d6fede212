#!/usr/bin/env bash
# round 2: exact_median option: parity subset + the two sweeps that found the median flips
set -u
O=gpurun_out/r2e2; mkdir -p $O
GSR_TEST_OPTS=exact_median=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "golden or live or oracle" > $O/pytest_exact.txt 2>&1; echo "pytest exit $?"; tail -n 2 $O/pytest_exact.txt
timeout 600 python tools/parity_fuzz.py --cases 800 --seed 2 --opt exact_median=1 --out $O/parity_fuzz_seed2_exact.txt > /dev/null 2>$O/fuzz.err; echo "fuzz exit $?"; tail -n 1 $O/parity_fuzz_seed2_exact.txt | cut -c1-250; grep "^FAIL" -A1 $O/parity_fuzz_seed2_exact.txt | cut -c1-220 | head -6
timeout 600 python tools/parity_fuzz.py --cases 900 --seed 3 --opt exact_median=1 --out $O/parity_fuzz_seed3_exact.txt > /dev/null 2>>$O/fuzz.err; echo "fuzz exit $?"; tail -n 1 $O/parity_fuzz_seed3_exact.txt | cut -c1-250; grep "^FAIL" -A1 $O/parity_fuzz_seed3_exact.txt | cut -c1-220 | head -6
bash tools/gpu_ab_opts.sh "C3 light" "exact_median=0" "exact_median=1" 2>&1 | tee $O/ab_exact.txt
tail -2 $O/fuzz.err
