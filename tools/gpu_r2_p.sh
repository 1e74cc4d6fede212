#!/usr/bin/env bash
# round 2: large randomised parity sweep against the reference build
set -u
O=gpurun_out/final2; mkdir -p $O
timeout 1700 python tools/parity_fuzz.py --cases 1500 --seed 2 --out $O/parity_fuzz_1500.txt > /dev/null 2>$O/parity_fuzz.err; echo "fuzz exit $?"; tail -n 1 $O/parity_fuzz_1500.txt | cut -c1-300; grep -c FAIL $O/parity_fuzz_1500.txt; tail -3 $O/parity_fuzz.err
