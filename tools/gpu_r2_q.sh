#!/usr/bin/env bash
# round 2: PDL + correctly rounded T restoration (-light): full GPU suite, two parity sweeps, bench
set -u
O=gpurun_out/r2q; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt; tail -n 4 $O/pytest_gpu.txt
timeout 900 python tools/parity_fuzz.py --cases 1500 --seed 2 --out $O/parity_fuzz_1500.txt > /dev/null 2>$O/parity_fuzz.err; echo "fuzz exit $?"; tail -n 1 $O/parity_fuzz_1500.txt | cut -c1-300; grep "^FAIL" -A2 $O/parity_fuzz_1500.txt | cut -c1-250 | head; tail -2 $O/parity_fuzz.err
timeout 900 python tools/parity_fuzz.py --cases 1500 --seed 3 --out $O/parity_fuzz_1500_seed3.txt > /dev/null 2>$O/parity_fuzz.err; echo "fuzz exit $?"; tail -n 1 $O/parity_fuzz_1500_seed3.txt | cut -c1-300; grep "^FAIL" -A2 $O/parity_fuzz_1500_seed3.txt | cut -c1-250 | head
bash tools/gpu_ab_opts.sh "C3 light;C4 light" "pdl=1" 2>&1 | tee $O/ab_light.txt
