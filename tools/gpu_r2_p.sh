#!/usr/bin/env bash
# round 2: randomised parity sweep against the reference build
set -u
O=gpurun_out/final2; mkdir -p $O
timeout 1500 python tools/parity_fuzz.py --cases 160 --seed 1 --out $O/parity_fuzz.txt > /dev/null 2>$O/parity_fuzz.err; echo "fuzz exit $?"; tail -n 2 $O/parity_fuzz.txt | cut -c1-300; grep -c FAIL $O/parity_fuzz.txt; tail -3 $O/parity_fuzz.err
